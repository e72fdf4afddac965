"""Small host utilities: a Log with the reference's method names, and a reader for the
subset of `[pyrat]` configuration keys that drive the opacity path."""
import configparser
import os
import string
import sys
import textwrap
from types import SimpleNamespace

import numpy as np

from . import constants as pc


class Log:
    """Stand-in for mc3.utils.Log (head/msg/debug/warning/error; error raises)."""

    def __init__(self, logname=None, verb=2):
        self.logname = logname
        self.verb = verb
        self.file = open(logname, 'w') if logname else None
        self.warnings = []

    def _emit(self, text, level, indent=0):
        if self.verb >= level:
            pad = ' ' * indent
            out = '\n'.join(pad + line for line in str(text).split('\n'))
            print(out)
            sys.stdout.flush()
            if self.file is not None:
                self.file.write(out + '\n')
                self.file.flush()

    def head(self, text, indent=0, **kw):
        self._emit(text, 1, indent)

    def msg(self, text, indent=0, **kw):
        self._emit(text, 2, indent)

    def debug(self, text, indent=0, **kw):
        self._emit(text, 3, indent)

    def warning(self, text, **kw):
        self.warnings.append(text)
        self._emit('Warning: ' + str(text), 1)

    def error(self, text, **kw):
        self._emit('Error: ' + str(text), 0)
        raise ValueError(text)

    def close(self):
        if self.file is not None:
            self.file.close()
            self.file = None


class Formatted_Write(string.Formatter):
    """Accumulate formatted, wrapped text (same contract as the reference's
    tools.Formatted_Write, tools/tools.py:736-829): `None` prints as 'None' under any format
    spec, NumPy arrays are rendered under temporary printoptions, lines wrap at 80 columns."""

    def __init__(self, indent=0, si=4, fmt=None, edge=None, lw=80, prec=None):
        self.text = ''
        self.indent = indent
        self.si = si
        self.fmt, self.edge, self.lw, self.prec = fmt, edge, lw, prec

    def format_field(self, value, spec):
        if value is None:
            return 'None'
        return super().format_field(value, spec)

    def write(self, text, *format, **numpy_fmt):
        opts = {'fmt': self.fmt, 'edge': self.edge, 'lw': self.lw, 'prec': self.prec}
        opts.update(numpy_fmt)
        printopts = {
            'formatter': opts['fmt'],
            'edgeitems': opts['edge'],
            'threshold': None if opts['edge'] is None else 2 * opts['edge'],
            'linewidth': opts['lw'],
            'precision': opts['prec'],
        }
        with np.printoptions(**printopts):
            text = super().format(text, *format)
        first = ' ' * self.indent
        rest = first if self.si is None else ' ' * self.si
        for line in text.splitlines():
            self.text += textwrap.fill(line, break_long_words=False, initial_indent=first,
                                       subsequent_indent=rest, width=80)
            self.text += '\n'


def _value_units(text, default_units=None):
    """'1.1 um' -> 1.1e-4 (CGS); a bare number uses default_units."""
    parts = text.split()
    val = float(parts[0])
    if len(parts) > 1:
        return val * pc.u(parts[1])
    if default_units is not None:
        return val * pc.u(default_units)
    return val


_FLOAT = ['wnlow', 'wnhigh', 'wnstep', 'resolution', 'tmin', 'tmax', 'tstep', 'ethresh',
          'voigt_extent', 'voigt_cutoff', 'voigt_dmin', 'voigt_dmax', 'voigt_lmin',
          'voigt_lmax', 'voigt_dlratio']
_INT = ['wnosamp', 'voigt_ndop', 'voigt_nlor', 'nlayers', 'ncpu', 'verb']
_DEFAULTS = dict(ethresh=1e-30, voigt_extent=300.0, voigt_cutoff=25.0, voigt_ndop=50,
                 voigt_nlor=100, voigt_dlratio=0.1, ncpu=1, verb=2)


def parse(cfile):
    """Read the `[pyrat]` keys relevant to runmode=opacity / LBL extinction
    (names and defaults of tools/parser.py:456-521,676-696,924-961,993-995)."""
    if not os.path.isfile(cfile):
        raise ValueError(f"Configuration file '{cfile}' does not exist")
    cp = configparser.ConfigParser()
    cp.optionxform = str
    cp.read(cfile)
    if 'pyrat' not in cp.sections():
        raise ValueError(f"\nInvalid configuration file: '{cfile}', no [pyrat] section")
    sec = cp['pyrat']
    args = SimpleNamespace(configfile=cfile)
    for key in _FLOAT:
        setattr(args, key, float(sec[key].split()[0]) if key in sec else _DEFAULTS.get(key))
    for key in _INT:
        setattr(args, key, int(sec[key].split()[0]) if key in sec else _DEFAULTS.get(key))
    wlunits = sec.get('wlunits', 'um')
    args.wlunits = wlunits
    for key in ['wl_low', 'wl_high', 'wlstep']:
        setattr(args, key, _value_units(sec[key], wlunits) if key in sec else None)
    for key in ['ptop', 'pbottom']:
        setattr(args, key, _value_units(sec[key], 'bar') / pc.bar if key in sec else None)
    for key in ['runmode', 'logfile', 'atmfile', 'single_isotope']:
        setattr(args, key, sec.get(key))
    args.tlifile = sec['tlifile'].split() if 'tlifile' in sec else None
    cs = sec.get('sampled_cross_sec', sec.get('extfile'))
    args.sampled_cs = cs.split() if cs is not None else None
    if args.sampled_cs is None and args.runmode == 'opacity' and args.logfile:
        args.sampled_cs = [os.path.splitext(args.logfile)[0] + '.npz']  # parser.py:695-696
    root = os.path.dirname(os.path.abspath(cfile))
    # runmode = tli (parser.py:456-460,709): databases, their types and partition functions
    args.dblist = sec['dblist'].split() if 'dblist' in sec else None
    args.dbtype = sec['dbtype'].split() if 'dbtype' in sec else None
    args.pflist = sec['pflist'].split() if 'pflist' in sec else None
    if args.dblist is not None:
        args.dblist = [p if os.path.isabs(p) else os.path.join(root, p) for p in args.dblist]
    if args.pflist is not None:
        args.pflist = [p if (p in ('tips', 'poly') or os.path.isabs(p))
                       else os.path.join(root, p) for p in args.pflist]
    if args.tlifile is None and args.runmode == 'tli' and sec.get('logfile'):
        args.tlifile = [os.path.splitext(sec.get('logfile'))[0] + '.tli']  # parser.py:693-694
    for key in ['logfile', 'atmfile']:
        v = getattr(args, key)
        if v is not None and not os.path.isabs(v):
            setattr(args, key, os.path.join(root, v))
    for key in ['tlifile', 'sampled_cs']:
        v = getattr(args, key)
        if v is not None:
            setattr(args, key, [p if os.path.isabs(p) else os.path.join(root, p) for p in v])
    return args
