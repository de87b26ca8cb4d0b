"""
Util functions (host side) — mirror of the reference's ``rtgs.utils`` package.
"""
