"""Algorithm registry with the reference's surface (semilearn/core/utils/registry.py:25-36):
`@ALGORITHMS.register('name')`, `ALGORITHMS['name']`, `'name' in ALGORITHMS`, `.keys()`.
This registry is the package's own.  `semireward_b200.integration.register_into_reference()` is the
explicit call that puts these classes, combined with the reference's AlgorithmBase loop, into `semilearn`'s registry."""
from __future__ import annotations


class Register:
    def __init__(self, registry_name: str):
        self._dict = {}
        self._name = registry_name

    def __setitem__(self, key, value):
        if not callable(value):
            raise Exception(f"Value of a Registry must be a callable!\nValue: {value}")
        if key is None:
            key = value.__name__
        self._dict[key] = value

    def register(self, target):
        def add(key, value):
            self[key] = value
            return value
        if callable(target):       # @reg.register
            return add(None, target)
        return lambda x: add(target, x)   # @reg.register('alias')

    def __getitem__(self, key):
        return self._dict[key]

    def __contains__(self, key):
        return key in self._dict

    def keys(self):
        return self._dict.keys()


ALGORITHMS = Register("algorithms")
