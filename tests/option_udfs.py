"""UDFs over Option[T] columns (rows with None in the normal case) for tests/test_option.py — kept in a file so that
inspect.getsource works for the front end."""


def is_none(x):
    return x['a'] is None


def eq_none(x):
    return x['a'] == None  # noqa: E711


def ne_none(x):
    return x['a'] != None  # noqa: E711


def is_not_none(x):
    return x['a'] is not None


def fill_zero(x):
    if x['a'] is None:
        return 0
    return x['a'] + 1


def fill_ifexp(x):
    return x['a'] * 2 if x['a'] is not None else -1


def use_raises(x):
    return x['a'] + x['b']          # TypeError on the rows where a is None


def passthrough(x):
    return (x['b'], x['a'], x['s'])


def maybe_none(x):
    if x['b'] > 50:
        return None
    return x['b'] * 3


def str_len_or_none(x):
    if x['s'] is None:
        return None
    return len(x['s'])


def eq_value(x):
    return x['a'] == 7


def ne_value(x):
    return x['a'] != x['b']


def both_option_eq(x):
    return x['a'] == x['c']


def truthy(x):
    if x['a']:
        return 1
    return 0


def str_truthy_lower(x):
    if x['s']:
        return x['s'].lower()
    return 'none-or-empty'


def and_guard(x):
    return x['a'] is not None and x['a'] > 3


def or_guard(x):
    return x['s'] is None or 'a' in x['s']


def none_in_tuple(x):
    return (x['b'], None if x['b'] % 2 == 0 else x['s'])


def str_of_option(x):
    return str(x['a'])              # 'None' for None rows: resolved by CPython after the device raises TypeError


def less_than(x):
    return x['a'] < 5               # TypeError for None


UDFS = [is_none, eq_none, ne_none, is_not_none, fill_zero, fill_ifexp, use_raises, passthrough, maybe_none, str_len_or_none, eq_value, ne_value,
        both_option_eq, truthy, str_truthy_lower, and_guard, or_guard, none_in_tuple, str_of_option, less_than]
