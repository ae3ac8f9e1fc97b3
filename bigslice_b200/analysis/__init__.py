"""analysis subpackage."""
