"""Placeholder module mirroring `uisrnn.contrib.contrib_template` of the reference."""


def example_function():
  """Returns True (the reference's template function)."""
  return True
