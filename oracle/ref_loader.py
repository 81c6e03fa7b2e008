"""Run the REAL reference code in the build container.  TEST INFRASTRUCTURE ONLY.

``/root/reference`` exists only in the build container (never on the GPU box),
so this module is used for exactly two things:

* ``oracle/make_golden.py`` - mint the golden vectors under ``tests/golden/``;
* ``tests/test_oracle_vs_reference.py`` - check the numpy restatement in
  ``oracle/fithic_oracle.py`` against the reference itself (skipped when
  ``/root/reference`` is absent).

No reference SOURCE is copied into the repository: ``fithic.py`` is read from
where it lies, patched in memory with the closed edit list below and exec'd;
``blueberry.pyx`` is cythonized from where it lies, verbatim, with the build
products going to the git-ignored ``oracle/_ref/``.

Closed edit list for ``blueberry/fithic.py`` (Python 2 -> 3, nothing else):
  :157        ``print "..."`` statement -> ``print("...")`` call
  :167 :298 :314 :316   ``/`` -> ``//``  (Python-2 integer division is part of
              the reference's semantics: first desiredPerBin, maxFrag,
              possibleIntraAllCount, possibleInterAllCount)
  :141 :243 :285 :409   ``gzip.open(..., 'r')`` -> ``'rt'``
  :410        ``gzip.open(..., 'w')`` -> ``'wt'``
plus a no-op ``matplotlib`` stub on ``sys.modules`` (matplotlib is not installed
and only draws the PNG).  The dangling ``main()`` at :489-490 never runs because
the module is exec'd under a name other than ``__main__``.
"""
import os
import subprocess
import sys
import types

REFERENCE_ROOT = os.environ.get("BLUEBERRY_REFERENCE", "/root/reference")
_HERE = os.path.dirname(os.path.abspath(__file__))
REF_BUILD_DIR = os.path.join(_HERE, "_ref")

_EDITS = [
    (157, 'print "Out of " + str(i+1) + " loci " +str(discarded) +" were discarded with biases not in range [0.5 2]\\n\\n"',
          'print("Out of " + str(i+1) + " loci " +str(discarded) +" were discarded with biases not in range [0.5 2]\\n\\n")'),
    (167, ")/n_bins", ")//n_bins"),
    (298, "resolution/2", "resolution//2"),
    (314, "(n*(n+1))/2", "(n*(n+1))//2"),
    (316, "possibleInterAllCount /= 2", "possibleInterAllCount //= 2"),
    (141, "'r')", "'rt')"),
    (243, "'r')", "'rt')"),
    (285, "'r')", "'rt')"),
    (409, "'r')", "'rt')"),
    (410, "'w')", "'wt')"),
]


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "blueberry", "fithic.py"))


class _Anything(object):
    """Absorbs any attribute access / call (stands in for matplotlib objects)."""

    def __getattr__(self, name):
        return _Anything()

    def __call__(self, *a, **k):
        return _Anything()


def _install_matplotlib_stub():
    if "matplotlib" in sys.modules and not getattr(sys.modules["matplotlib"], "_bbk_stub", False):
        return
    mpl = types.ModuleType("matplotlib")
    mpl._bbk_stub = True
    mpl.use = lambda *a, **k: None
    plt = types.ModuleType("matplotlib.pyplot")
    plt.__getattr__ = lambda name: _Anything()
    mpl.pyplot = plt
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = plt


def load_reference_fithic():
    """Return a FRESH module object holding the patched reference fithic.py.

    The reference keeps all totals in module globals that are never reset
    (fithic.py:25-42), so every oracle run must use a fresh module.
    """
    path = os.path.join(REFERENCE_ROOT, "blueberry", "fithic.py")
    with open(path) as fh:
        lines = fh.read().split("\n")
    for lineno, old, new in _EDITS:
        line = lines[lineno - 1]
        if old not in line:
            raise RuntimeError("reference fithic.py:%d changed; edit list no longer applies: %r" % (lineno, line))
        lines[lineno - 1] = line.replace(old, new)
    _install_matplotlib_stub()
    mod = types.ModuleType("blueberry_reference_fithic")
    mod.__file__ = path
    exec(compile("\n".join(lines), path, "exec"), mod.__dict__)
    return mod


_SETUP = r'''
import os, sys
from setuptools import setup, Extension
from Cython.Build import cythonize
import numpy
src = sys.argv.pop(1)
ext = Extension("blueberry_ref.blueberry", [src], include_dirs=[numpy.get_include()],
                define_macros=[("NPY_NO_DEPRECATED_API", "NPY_1_7_API_VERSION")])
setup(name="blueberry_ref", ext_modules=cythonize([ext], language_level=2, build_dir="cy_build", quiet=True),
      script_args=["build_ext", "--build-lib", ".", "--build-temp", "cy_tmp", "-q"])
'''


def build_reference_cython(force=False):
    """Cythonize /root/reference/blueberry/blueberry.pyx VERBATIM into oracle/_ref/.

    blueberry.pyx star-imports its sibling ``utils`` (blueberry.pyx:15) whose real
    source is Python-2 syntax; the stub written here holds only the four
    constants of utils.py:23-26.  Returns the import directory, or None when the
    reference is absent and nothing was prebuilt.
    """
    pkg = os.path.join(REF_BUILD_DIR, "blueberry_ref")
    have = os.path.isdir(pkg) and any(f.startswith("blueberry.") and f.endswith(".so") for f in os.listdir(pkg))
    if have and not force:
        return REF_BUILD_DIR
    if not reference_available():
        return None
    os.makedirs(pkg, exist_ok=True)
    with open(os.path.join(pkg, "__init__.py"), "w") as fh:
        fh.write("# build product of oracle/ref_loader.py (git-ignored)\n")
    with open(os.path.join(pkg, "utils.py"), "w") as fh:
        fh.write("# stub for blueberry/utils.py:23-26 (the four constants blueberry.pyx needs)\n"
                 "Q_LOWER_BOUND = 0.01\nQ_UPPER_BOUND = 0.50\n"
                 "HIGH_FITHIC_CUTOFF = 10000000\nLOW_FITHIC_CUTOFF = 25000\n")
    with open(os.path.join(REF_BUILD_DIR, "_setup_ref.py"), "w") as fh:
        fh.write(_SETUP)
    # Cython wants the .pyx inside the package tree it is compiled for; a symlink keeps
    # the source where it lies.
    link = os.path.join(pkg, "blueberry.pyx")
    if os.path.lexists(link):
        os.remove(link)
    os.symlink(os.path.join(REFERENCE_ROOT, "blueberry", "blueberry.pyx"), link)
    try:
        subprocess.check_call([sys.executable, "_setup_ref.py", os.path.join("blueberry_ref", "blueberry.pyx")],
                              cwd=REF_BUILD_DIR, stdout=subprocess.DEVNULL)
    finally:
        os.remove(link)  # the symlink would dangle on the GPU box
    return REF_BUILD_DIR


def load_reference_cython():
    """Import the verbatim-compiled blueberry.pyx (benjamini_hochberg, count_band_regions)."""
    d = build_reference_cython()
    if d is None:
        return None
    if d not in sys.path:
        sys.path.insert(0, d)
    import importlib
    return importlib.import_module("blueberry_ref.blueberry")


_SETUP_DT = r'''
import sys
from setuptools import setup, Extension
from Cython.Build import cythonize
import numpy
ext = Extension("blueberry_ref.datatypes", ["blueberry_ref/datatypes.pyx"], include_dirs=[numpy.get_include()],
                define_macros=[("NPY_NO_DEPRECATED_API", "NPY_1_7_API_VERSION")])
setup(name="blueberry_ref_dt", ext_modules=cythonize([ext], language_level=2, build_dir="cy_build", quiet=True),
      script_args=["build_ext", "--build-lib", ".", "--build-temp", "cy_tmp", "-q"])
'''


def build_reference_datatypes(force=False):
    """Cythonize /root/reference/blueberry/datatypes.pyx VERBATIM into oracle/_ref/ (next to blueberry.pyx's build).
    Its constructors read hard-coded NFS path templates held in module globals (datatypes.pyx:25-29);
    load_reference_datatypes() points those globals at temporary files, which is how ContactMap is driven here."""
    if build_reference_cython(force) is None:
        return None
    pkg = os.path.join(REF_BUILD_DIR, "blueberry_ref")
    have = any(f.startswith("datatypes.") and f.endswith(".so") for f in os.listdir(pkg))
    if have and not force:
        return REF_BUILD_DIR
    if not reference_available():
        return None
    with open(os.path.join(REF_BUILD_DIR, "_setup_ref_dt.py"), "w") as fh:
        fh.write(_SETUP_DT)
    link = os.path.join(pkg, "datatypes.pyx")
    if os.path.lexists(link):
        os.remove(link)
    os.symlink(os.path.join(REFERENCE_ROOT, "blueberry", "datatypes.pyx"), link)
    try:
        subprocess.check_call([sys.executable, "_setup_ref_dt.py"], cwd=REF_BUILD_DIR, stdout=subprocess.DEVNULL)
    finally:
        os.remove(link)
    return REF_BUILD_DIR


def load_reference_datatypes():
    """Import the verbatim-compiled datatypes.pyx.  It does `from blueberry import *` at import (datatypes.pyx:23): the
    compiled blueberry.pyx is registered under that name first."""
    d = build_reference_datatypes()
    if d is None:
        return None
    bb = load_reference_cython()
    sys.modules.setdefault("blueberry", bb)
    import importlib
    return importlib.import_module("blueberry_ref.datatypes")
