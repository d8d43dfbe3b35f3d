"""Mirror of codes/SRN/options (only the files the SRN hot path needs)."""


def _see_reference_package(name):
    """When this mirror shadows the reference package of the same name (dasr_b200.launch / PYTHONPATH overlay), keep the
    reference's OTHER submodules importable (e.g. utils.receptive_cal, which test.py imports): append the shadowed
    directory to this package's search path — files that exist here still win."""
    import os
    import sys
    here = [os.path.abspath(p) for p in __path__]
    for d in sys.path:
        cand = os.path.abspath(os.path.join(d or '.', name))
        if os.path.isdir(cand) and cand not in here and os.path.exists(os.path.join(cand, '__init__.py')):
            __path__.append(cand)


if __name__ == 'options':          # imported as the top-level package, i.e. as the drop-in
    _see_reference_package('options')
