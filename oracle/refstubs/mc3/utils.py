"""Minimal Log object with the methods the reference calls on tli/opacity paths."""
import sys


class Log:
    def __init__(self, logname=None, verb=2, append=False, width=70):
        self.logname = logname
        self.verb = verb
        self.width = width
        self.sep = 70 * ":"
        self.file = None
        self.warnings = []
        if logname is not None:
            self.file = open(logname, "a" if append else "w")

    def write(self, text):
        if self.file is not None:
            self.file.write(text + "\n")
            self.file.flush()
        print(text)
        sys.stdout.flush()

    def _emit(self, text, level, indent=0):
        if self.verb >= level:
            pad = " " * indent
            self.write("\n".join(pad + line for line in str(text).split("\n")))

    def head(self, text, verb=1, indent=0, **kw):
        self._emit(text, verb, indent)

    def msg(self, text, verb=2, indent=0, **kw):
        self._emit(text, verb, indent)

    def debug(self, text, verb=3, indent=0, **kw):
        self._emit(text, verb, indent)

    def progress(self, text, verb=2, indent=0, **kw):
        self._emit(text, verb, indent)

    def warning(self, text, **kw):
        self.warnings.append(text)
        self._emit("Warning: " + str(text), 1)

    def error(self, text, tracklev=-1, **kw):
        self._emit("Error: " + str(text), 0)
        raise RuntimeError(text)

    def close(self):
        if self.file is not None:
            self.file.close()
            self.file = None


def burn(*args, **kwargs):
    raise NotImplementedError("mc3 stub")
