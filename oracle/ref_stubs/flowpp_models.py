"""Empty stub for the module the reference imports but does not ship (utils.py:11)."""
