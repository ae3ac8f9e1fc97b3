"""slicetest subpackage."""
