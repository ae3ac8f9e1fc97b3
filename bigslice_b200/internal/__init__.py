"""internal subpackage."""
