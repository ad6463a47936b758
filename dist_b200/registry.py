"""Name -> class tables that drive model construction.

Mirrors the contract of the reference's ``utils/registry.py:6-65``: ``register()`` is a
decorator factory that keys an entry by its ``__name__``, registering a name twice is an
assertion error, and ``get`` answers ``None`` for unknown names (``models/base/builder.py:30``
relies on that to fall back to ``BaseVideoModel``).
"""


class Registry:
    def __init__(self, table_name=""):
        self.table_name = table_name
        self._entries = {}

    def register(self):
        def _add(obj):
            key = obj.__name__
            assert isinstance(key, str)
            assert key not in self._entries, "{} {} already registered.".format(self.table_name, key)
            self._entries[key] = obj
            return obj

        return _add

    def get(self, name):
        return self._entries.get(name, None)

    def get_all_registered(self):
        return self._entries.keys()

    def __contains__(self, name):
        return name in self._entries

    def __repr__(self):
        return "Registry({!r}: {})".format(self.table_name, sorted(self._entries))
