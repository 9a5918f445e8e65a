"""Import stub for the un-vendored `nerv` package (v0.4.0).

Only used by tools/make_golden.py IN THE BUILD CONTAINER to import the
reference modules from /root/reference and generate golden vectors.
Holds no hot-path arithmetic.  Never imported by the product or the tests.
"""
