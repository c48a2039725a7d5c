"""Minimal levelled logger with the two entry points the reference uses from the third-party
`colortimelog` package: `.print(level, message)` (uisrnn.py:166,299,312,345) and `.info(message)`
(tests/integration_test.py:112)."""
import sys
import time

_LEVEL_NAMES = {0: 'FATAL', 1: 'ERROR', 2: 'WARN', 3: 'INFO'}


class Logger:
  def __init__(self, verbosity=3, stream=None):
    self.verbosity = verbosity
    self._stream = stream

  def print(self, level, message):
    """Emits `message` when `level` <= verbosity."""
    if level > self.verbosity:
      return
    stream = self._stream or sys.stdout
    stream.write('[{} {}] {}\n'.format(time.strftime('%Y-%m-%d %H:%M:%S'),
                                      _LEVEL_NAMES.get(level, 'DEBUG'), message))
    stream.flush()

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
