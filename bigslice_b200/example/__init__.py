"""example subpackage."""
