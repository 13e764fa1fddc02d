"""INI configuration with the accessor names the reference's solvers use.

Same call surface as ``pyfr/inifile.py:31-161`` (``get`` with a default that
is written back, ``getint/getfloat/getbool``, ``items_as``, ``getexpr``) so
host code and backends can be handed either object.
"""

from configparser import ConfigParser, NoOptionError, NoSectionError
import re

_missing = object()
_truth = {'1': True, 'yes': True, 'true': True, 'on': True,
          '0': False, 'no': False, 'false': False, 'off': False}


def _floatify(m):
    tok = m[0]
    return tok if any(ch in tok for ch in '.eE') else tok + '.'


class Config:
    def __init__(self, text=None):
        self._cp = ConfigParser(inline_comment_prefixes=[';', '#'])
        self._cp.optionxform = str

        if text:
            self._cp.read_string(text)

    def set(self, section, option, value):
        if not self._cp.has_section(section):
            self._cp.add_section(section)

        self._cp.set(section, option, str(value))

    def hasopt(self, section, option):
        return self._cp.has_option(section, option)

    def get(self, section, option, default=_missing):
        try:
            return self._cp.get(section, option)
        except (NoSectionError, NoOptionError):
            if default is _missing:
                raise

            self.set(section, option, default)
            return self._cp.get(section, option)

    def getint(self, section, option, default=_missing):
        return int(self.get(section, option, default))

    def getfloat(self, section, option, default=_missing):
        return float(self.get(section, option, default))

    def getbool(self, section, option, default=_missing):
        return _truth[self.get(section, option, default).lower()]

    def getexpr(self, section, option, default=_missing, subs={}):
        expr = self.get(section, option, default)

        if not re.match(r'[A-Za-z0-9_ \t\n\r.,+\-*/%()]+$', expr):
            raise ValueError('Invalid characters in expression')

        if subs:
            names = '|'.join(map(re.escape, subs))
            expr = re.sub(rf'\b({names})\b', lambda m: str(subs[m[1]]), expr)

        # Promote integer literals so C and Python agree on division
        expr = re.sub(r'\b((\d+\.?\d*)|(\.\d+))([eE][+-]?\d+)?(?![^[]*\])',
                      _floatify, expr)

        return f'({expr})'

    def items(self, section, prefix=''):
        return self.items_as(section, str, prefix)

    def items_as(self, section, conv, prefix=''):
        out = {}

        if self._cp.has_section(section):
            for key, val in self._cp.items(section):
                if key.startswith(prefix):
                    try:
                        out[key] = conv(val)
                    except ValueError:
                        pass

        return out

    def sections(self):
        return self._cp.sections()
