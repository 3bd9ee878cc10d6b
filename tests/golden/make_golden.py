"""Generate tests/golden/golden.npz by running the UNMODIFIED reference
(/root/reference/src/tools21cm/power_spectrum.py) on the seeded case matrix in
cases.py.  Runs only in the build container (the reference tree does not travel
to the GPU box); the resulting .npz is committed.

    python tests/golden/make_golden.py

The reference package ``__init__`` needs astropy/skimage (absent here), so the
hot-path module is imported through a stub package.
"""
import os
import sys
import time
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, HERE)

from cases import CASES, build_inputs, build_kwargs  # noqa: E402

REF_DIR = "/root/reference/src/tools21cm"


def load_reference():
    pkg = types.ModuleType("tools21cm")
    pkg.__path__ = [REF_DIR]
    sys.modules["tools21cm"] = pkg
    import tools21cm.power_spectrum as ref  # noqa
    # power_legendre.py:47 calls _get_mu with three arguments (it takes four, power_spectrum.py:481): the module is
    # imported unmodified and only that name is rebound to a three-argument shim, so that the function can run at all
    import tools21cm.power_legendre as leg  # noqa
    leg._get_mu = lambda k_comp, k, los_axis: ref._get_mu(k_comp, k, los_axis, True)
    ref.power_spectrum_multipoles = leg.power_spectrum_multipoles
    import tools21cm.position_dependent_power_spectrum as pdp  # noqa
    for name in ("integrated_bispectrum", "integrated_bispectrum_cross", "power_spectrum_response"):
        setattr(ref, name, getattr(pdp, name))
    return ref


def main():
    ref = load_reference()
    out = {}
    for case in CASES:
        t0 = time.time()
        fn = getattr(ref, case["fn"])
        arrays = build_inputs(case)
        kw = build_kwargs(case)
        if case["fn"] in ("power_spectrum_1d", "cross_power_spectrum_1d",
                          "power_spectrum_mu", "cross_power_spectrum_mu"):
            kw["return_n_modes"] = True
        with np.errstate(all="ignore"):  # the package __init__ does numpy.seterr(all='ignore')
            res = fn(*arrays, **kw)
        if not isinstance(res, tuple):
            res = (res,)
        if case["fn"] == "power_spectrum_multipoles":      # (dict, k) -> P0, P2, P4, nmodes, k
            res = (res[0]["P0"], res[0]["P2"], res[0]["P4"], res[0]["nmodes"], res[1])
        for j, r in enumerate(res):
            out["%s::%d" % (case["id"], j)] = np.asarray(r)
        print("%-32s %6.2fs  %s" % (case["id"], time.time() - t0,
                                    [np.asarray(r).shape for r in res]), flush=True)
    path = os.path.join(HERE, "golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
