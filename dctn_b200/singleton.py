"""Metaclass giving one shared instance per class (API of dctn/singleton.py:1-7)."""


class Singleton(type):
    _instances = {}

    def __call__(cls, *args, **kwargs):
        inst = Singleton._instances.get(cls)
        if inst is None:
            inst = Singleton._instances[cls] = super().__call__(*args, **kwargs)
        return inst
