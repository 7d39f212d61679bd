"""Import stub so the read-only reference (flow_models/wolf/*) imports here.
Test infrastructure only; changes no arithmetic."""


def overrides(f):
    return f
