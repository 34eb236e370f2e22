"""Argument metadata with the reference's record (torchrec/utils/argument/ArgumentDescription.py:20-80): the models'
``get_argument_descriptions()`` lists return these, so a caller that builds its command line from them (the
reference's ``IWithArguments``) finds the same fields.  Only the record and its validation live here — the argparse
plumbing belongs to the reference's task layer, which is out of this package's scope."""
from typing import Any, List, Optional, Type


class ArgumentDescription:
    _TYPES = (str, int, float, bool)

    def __init__(self, name: str, type_: Type, help_info: str, is_logged: bool = True, default_value: Any = None,
                 legal_value_list: Optional[List[Any]] = None, lower_open_bound=None, lower_closed_bound=None,
                 upper_open_bound=None, upper_closed_bound=None):
        if type_ not in self._TYPES:
            raise ValueError(f"argument {name}: type must be one of str, int, float, bool")
        if default_value is not None and not isinstance(default_value, type_):
            raise ValueError(f"argument {name}: default {default_value!r} is not a {type_.__name__}")
        if legal_value_list:
            lower_open_bound = lower_closed_bound = upper_open_bound = upper_closed_bound = None
        elif any(b is not None for b in (lower_open_bound, lower_closed_bound, upper_open_bound, upper_closed_bound)):
            if type_ not in (int, float):
                raise ValueError(f"argument {name}: bounds need a numeric type")
        self.name, self.type, self.help_info, self.is_logged = name, type_, help_info, is_logged
        self.default_value, self.legal_value_list = default_value, legal_value_list
        self.lower_open_bound, self.lower_closed_bound = lower_open_bound, lower_closed_bound
        self.upper_open_bound, self.upper_closed_bound = upper_open_bound, upper_closed_bound

    def check(self, value: Any) -> None:
        """Raise ValueError if ``value`` violates the description (type, legal values, bounds)."""
        if not isinstance(value, self.type) or (self.type is not bool and isinstance(value, bool)):
            raise ValueError(f"argument {self.name}: expected {self.type.__name__}, got {value!r}")
        if self.legal_value_list and value not in self.legal_value_list:
            raise ValueError(f"argument {self.name}: {value!r} not in {self.legal_value_list}")
        if self.lower_open_bound is not None and not value > self.lower_open_bound:
            raise ValueError(f"argument {self.name}: {value!r} must be > {self.lower_open_bound}")
        if self.lower_closed_bound is not None and not value >= self.lower_closed_bound:
            raise ValueError(f"argument {self.name}: {value!r} must be >= {self.lower_closed_bound}")
        if self.upper_open_bound is not None and not value < self.upper_open_bound:
            raise ValueError(f"argument {self.name}: {value!r} must be < {self.upper_open_bound}")
        if self.upper_closed_bound is not None and not value <= self.upper_closed_bound:
            raise ValueError(f"argument {self.name}: {value!r} must be <= {self.upper_closed_bound}")
