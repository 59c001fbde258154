"""Import-path mirror of the reference package `flex_gym` (setup.py:78-88) over the B200 implementation."""
