"""cmd subpackage."""
