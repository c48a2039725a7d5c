"""Stand-in for the third-party `colortimelog` package (not installed, no network).

TEST INFRASTRUCTURE ONLY.  The reference imports `colortimelog` purely for logging
(`/root/reference/uisrnn/uisrnn.py:16,107` constructs `Logger(verbosity)`; call sites are
`.print(level, msg)` at `:166,299,312,345` and `.info(msg)` in
`/root/reference/tests/integration_test.py:112`).  No arithmetic goes through it, so this shim
cannot affect parity.  It is put on `sys.path` only by `oracle/` scripts that import the
reference (golden generation here, reference timing in `bench.py --impl reference`).
"""
import sys
import time


class Logger:
  def __init__(self, verbosity=3, stream=None):
    self.verbosity = verbosity
    self.stream = stream or sys.stderr

  def print(self, level, message):
    if level <= self.verbosity:
      self.stream.write('[{}] {}\n'.format(time.strftime('%H:%M:%S'), message))

  def fatal(self, message):
    self.print(0, message)

  def error(self, message):
    self.print(1, message)

  def warning(self, message):
    self.print(2, message)

  def info(self, message):
    self.print(3, message)

  def debug(self, message):
    self.print(4, message)
