"""tenpy shim subpackage (see tenpy/__init__.py)."""
