"""Name -> class registry with the semantics of fvcore.common.registry.Registry
(config_slowfast/fvcore/fvcore/common/registry.py:40-72): `register()` works as a decorator or a call, keys are
`cls.__name__`, duplicates assert, `get()` raises KeyError for unknown names."""


class Registry:
    def __init__(self, name):
        self._name = name
        self._obj_map = {}

    def _do_register(self, name, obj):
        assert name not in self._obj_map, "An object named '{}' was already registered in '{}' registry!".format(
            name, self._name)
        self._obj_map[name] = obj

    def register(self, obj=None, name=None):
        if obj is None:
            def deco(func_or_class):
                self._do_register(name or func_or_class.__name__, func_or_class)
                return func_or_class
            return deco
        self._do_register(name or obj.__name__, obj)
        return obj

    def get(self, name):
        ret = self._obj_map.get(name)
        if ret is None:
            raise KeyError("No object named '{}' found in '{}' registry!".format(name, self._name))
        return ret

    def __contains__(self, name):
        return name in self._obj_map

    def names(self):
        return sorted(self._obj_map)
