"""Minimal astropy stand-in so the unmodified reference imports in the build container
(astropy is not installed and there is no network).  Used ONLY by make_golden.py."""
