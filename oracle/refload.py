"""
TEST INFRASTRUCTURE -- loader for the reference's own pure-NumPy modules.

Loads, by file path and *unmodified*, the parts of /root/reference that have no
native dependency, behind ~40 lines of stub modules for the packages that are
absent from this image (mpi4py, pmesh, dask ...).  It exists for two jobs only:

  1. pinning the oracle restatement (oracle/*.py) against the reference code;
  2. generating the golden vectors committed under tests/golden/
     (tests/golden/make_golden.py).

/root/reference does not exist on the GPU box, so nothing that runs there may
import this module; `available()` tells a test whether it can be used.

Reference modules loaded (all verbatim):
  nbodykit/meshtools.py            -> SlabIterator, MeshSlab
  nbodykit/binned_statistic.py     -> BinnedStatistic
  nbodykit/algorithms/fftpower.py  -> project_to_basis, _find_unique_edges
  nbodykit/source/mesh/catalog.py  -> Compensate*, get_compensation
  nbodykit/mpirng.py               -> MPIRandomState
"""
import importlib.util
import os
import sys
import types

REF = os.environ.get("NBK_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF, "nbodykit"))


class FakeComm(object):
    """single-rank stand-in for mpi4py's COMM_WORLD (only what the loaded code calls)"""
    rank = 0
    size = 1

    def allgather(self, x):
        return [x]

    def allreduce(self, x, op=None):
        return x

    def bcast(self, x, root=0):
        return x

    def alltoall(self, x):
        return list(x)

    def Barrier(self):
        pass


_loaded = {}


def _stub(name, **members):
    m = types.ModuleType(name)
    m.__dict__.update(members)
    sys.modules[name] = m
    return m


def _load(modname, relpath):
    spec = importlib.util.spec_from_file_location(modname, os.path.join(REF, relpath))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[modname] = mod
    spec.loader.exec_module(mod)
    return mod


def load():
    """returns a namespace with the reference's functions; idempotent"""
    if _loaded:
        return _loaded["ns"]
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF)
    if "nbodykit" in sys.modules and not getattr(sys.modules["nbodykit"], "_oracle_stub", False):
        raise RuntimeError("a real 'nbodykit' is already imported; refusing to shadow it")

    import warnings
    warnings.filterwarnings("ignore", category=SyntaxWarning)

    # ---- stubs for absent third-party packages
    MPI = types.SimpleNamespace(COMM_WORLD=FakeComm(), MIN="min", LOR="lor", SUM="sum")
    _stub("mpi4py", MPI=MPI)
    _stub("mpi4py.MPI", **MPI.__dict__)

    class _Field(object):
        pass

    class _Resampler(object):
        def __init__(self, support):
            self.support = support

    pm = _stub("pmesh.pm", ParticleMesh=object, RealField=_Field, ComplexField=_Field,
               BaseComplexField=_Field, Field=_Field)
    window = _stub("pmesh.window", methods=dict(cic=_Resampler(2), tsc=_Resampler(3),
                                                pcs=_Resampler(4), nnb=_Resampler(1)))
    _stub("pmesh", pm=pm, window=window)

    # ---- a shell of the nbodykit package (its real __init__ needs mpi4py + dask)
    class CurrentMPIComm(object):
        @staticmethod
        def enable(func):
            return func

        @staticmethod
        def get():
            return MPI.COMM_WORLD

    nb = _stub("nbodykit", CurrentMPIComm=CurrentMPIComm, _oracle_stub=True,
               _global_options={"paint_chunk_size": 4 * 1024 * 1024,
                                "dask_chunk_size": 100000, "global_cache_size": 1e8})
    nb.__path__ = []

    def FrontPadArray(array, front, comm):
        # single rank: nothing in front of rank 0 (reference utils.py:350-370 reduces to this)
        assert front == 0
        return array

    _stub("nbodykit.utils", FrontPadArray=FrontPadArray)
    _stub("nbodykit.base")
    _stub("nbodykit.base.catalog", CatalogSourceBase=type("CatalogSourceBase", (), {}))
    _stub("nbodykit.base.mesh", MeshSource=type("MeshSource", (), {"actions": property(lambda s: [])}))
    _stub("nbodykit.source")
    _stub("nbodykit.source.mesh", FieldMesh=object)
    _stub("nbodykit.algorithms")

    ns = types.SimpleNamespace()
    ns.FakeComm = FakeComm
    ns.meshtools = _load("nbodykit.meshtools", "nbodykit/meshtools.py")
    ns.binned_statistic = _load("nbodykit.binned_statistic", "nbodykit/binned_statistic.py")
    ns.fftpower = _load("nbodykit.algorithms.fftpower", "nbodykit/algorithms/fftpower.py")
    ns.catalogmesh = _load("nbodykit.source.mesh.catalog", "nbodykit/source/mesh/catalog.py")
    ns.mpirng = _load("nbodykit.mpirng", "nbodykit/mpirng.py")

    ns.project_to_basis = ns.fftpower.project_to_basis
    ns.find_unique_edges = ns.fftpower._find_unique_edges
    ns.SlabIterator = ns.meshtools.SlabIterator
    ns.BinnedStatistic = ns.binned_statistic.BinnedStatistic
    ns.MPIRandomState = ns.mpirng.MPIRandomState
    ns.get_compensation = ns.catalogmesh.get_compensation
    _loaded["ns"] = ns
    return ns


class RefComplexField(object):
    """the minimum duck type `project_to_basis` needs (fftpower.py:570-572,645):
    `.x` (3 broadcastable coordinate arrays), `.pm.comm`, `.compressed`, `.dtype`, `[index]`"""

    def __init__(self, value, x, compressed=True):
        self.value = value
        self.x = x
        self.compressed = compressed
        self.dtype = value.dtype
        self.pm = types.SimpleNamespace(comm=FakeComm())

    def __getitem__(self, idx):
        return self.value[idx]
