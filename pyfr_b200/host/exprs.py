"""Evaluation of configuration-file expressions on NumPy arrays.

Equivalent in effect to the reference's ``npeval``
(``pyfr/nputil.py:107-121``): a restricted ``eval`` with a fixed
vocabulary of elementary functions.
"""

import re

import numpy as np

_vocab = {
    '__builtins__': {},
    'exp': np.exp, 'log': np.log, 'sin': np.sin, 'asin': np.arcsin,
    'cos': np.cos, 'acos': np.arccos, 'tan': np.tan, 'atan': np.arctan,
    'atan2': np.arctan2, 'abs': np.abs, 'pow': np.power, 'sqrt': np.sqrt,
    'tanh': np.tanh, 'pi': np.pi, 'max': np.maximum, 'min': np.minimum
}


def npeval(expr, names):
    if '^' in expr or '**' in expr:
        raise ValueError('Direct exponentiation is not supported; use pow')

    if not re.match(r'[A-Za-z0-9_ \t\n\r.,+\-*/%()]+$', expr):
        raise ValueError('Invalid characters in expression')

    known = '|'.join([*_vocab, *names])
    if re.search(rf'({known}|\))\s*\.', expr):
        raise ValueError('Invalid expression')

    return eval(expr, _vocab, dict(names))
