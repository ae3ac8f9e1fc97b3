"""archive subpackage."""
