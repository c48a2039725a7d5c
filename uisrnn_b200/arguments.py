"""Command-line / namespace configuration, mirroring the reference's three argument groups
(`/root/reference/uisrnn/arguments.py:30-205`): same flag names, short options, types, defaults
and return shape `(model_args, training_args, inference_args)`, so code written against
google/uis-rnn runs unchanged.  The flags are declared as data and materialised into argparse
parsers on demand.
"""
import argparse

_TRUE = frozenset(('yes', 'true', 't', 'y', '1'))
_FALSE = frozenset(('no', 'false', 'f', 'n', '0'))


def str2bool(value):
  """Parses a boolean flag value (arguments.py:21-27 of the reference)."""
  lowered = value.lower()
  if lowered in _TRUE:
    return True
  if lowered in _FALSE:
    return False
  raise argparse.ArgumentTypeError('Boolean value expected.')


# (flags, default, type, choices, help)
MODEL_FLAGS = (
    (('--observation_dim',), 256, int, None, 'Dimension of the observations (d-vectors).'),
    (('--rnn_hidden_size',), 512, int, None, 'Hidden units of each GRU layer.'),
    (('--rnn_depth',), 1, int, None, 'Number of stacked GRU layers.'),
    (('--rnn_dropout',), 0.2, float, None, 'Dropout between GRU layers (only used when depth >= 2).'),
    (('--transition_bias',), None, float, None,
     'p0 of Eq. (6) of the paper; estimated from the training labels (Eq. 13) when omitted.'),
    (('--crp_alpha',), 1.0, float, None, 'alpha of the distance-dependent CRP, Eq. (7); must be given.'),
    (('--sigma2',), None, float, None, 'sigma^2 of Eq. (11); learned from data when omitted.'),
    (('--verbosity',), 3, int, None,
     'Logging verbosity: 0 fatal, 1 error, 2 important steps, 3 all steps, >=4 debug.'),
    (('--enable_cuda',), True, str2bool, None, 'Run on cuda:0 when a CUDA device is available.'),
)
TRAINING_FLAGS = (
    (('--optimizer', '-o'), 'adam', None, ('adam',), 'Optimizer used by fit().'),
    (('--learning_rate', '-l'), 1e-3, float, None, 'Learning rate.'),
    (('--train_iteration', '-t'), 20000, int, None, 'Number of training iterations.'),
    (('--batch_size', '-b'), 10, int, None, 'Sub-sequences per training batch.'),
    (('--num_permutations',), 10, int, None, 'Block-permuted copies sampled per speaker sequence.'),
    (('--sigma_alpha',), 1.0, float, None, 'Inverse-gamma shape of the sigma^2 prior.'),
    (('--sigma_beta',), 1.0, float, None, 'Inverse-gamma scale of the sigma^2 prior.'),
    (('--regularization_weight', '-r'), 1e-5, float, None, 'Weight of the parameter-norm regulariser.'),
    (('--grad_max_norm',), 5.0, float, None, 'Gradient clipping norm.'),
    (('--enforce_cluster_id_uniqueness',), True, str2bool, None,
     'Prefix labels with a per-sequence random id when fit() receives a list, so equal labels '
     'in different sequences denote different speakers.'),
)
INFERENCE_FLAGS = (
    (('--beam_size', '-s'), 10, int, None, 'Beam width of the decoder.'),
    (('--look_ahead',), 1, int, None, 'Frames decided jointly per beam step.'),
    (('--test_iteration',), 2, int, None,
     'The test sequence is tiled this many times; labels of the last copy are returned.'),
)


def _make_parser(description, flags):
  parser = argparse.ArgumentParser(description=description, add_help=False)
  for names, default, typ, choices, text in flags:
    kwargs = {'default': default, 'help': text}
    if typ is not None:
      kwargs['type'] = typ
    if choices is not None:
      kwargs['choices'] = list(choices)
    parser.add_argument(*names, **kwargs)
  return parser


def parse_arguments(argv=None):
  """Returns `(model_args, training_args, inference_args)` namespaces.

  Unknown flags are rejected (a combined parser validates first), then each group picks its
  own flags with `parse_known_args` -- the behaviour of arguments.py:195-205 in the reference.
  """
  groups = (_make_parser('Model configurations.', MODEL_FLAGS),
            _make_parser('Training configurations.', TRAINING_FLAGS),
            _make_parser('Inference configurations.', INFERENCE_FLAGS))
  argparse.ArgumentParser(parents=list(groups)).parse_args(argv)
  return tuple(group.parse_known_args(argv)[0] for group in groups)
