"""Minimal stand-in for ml_collections.ConfigDict (attribute-style dict).
Test infrastructure only: lets the reference's configs/*.py be imported."""


class ConfigDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def lock(self):
        return self
