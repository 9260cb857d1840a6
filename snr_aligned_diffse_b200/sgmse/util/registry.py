"""String-keyed class tables used to pick backbones, SDEs, predictors and correctors by name.

Behavioural contract kept from the reference (sgmse-bbed/sgmse/util/registry.py:5-34): `Registry(kind)`,
`@reg.register(key)` as a class decorator (a second registration of a key warns and wins),
`reg.get_by_name(key)` raising `ValueError("<kind> with name '<key>' unknown.")`, `reg.get_all_names()`.
"""
import warnings


class Registry:
    def __init__(self, managed_thing):
        self.managed_thing = managed_thing      # used in messages only
        self._table = {}

    def _add(self, key, cls):
        if key in self._table:
            warnings.warn("%s with name '%s' doubly registered, old class will be replaced." % (self.managed_thing, key))
        self._table[key] = cls
        return cls

    def register(self, name):
        return lambda cls: self._add(name, cls)

    def get_by_name(self, name):
        try:
            return self._table[name]
        except KeyError:
            raise ValueError("%s with name '%s' unknown." % (self.managed_thing, name)) from None

    def get_all_names(self):
        return [*self._table]

    def __contains__(self, name):
        return name in self._table
