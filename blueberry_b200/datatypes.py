"""FithicContactMap - the consumer of the pass's output file (blueberry/datatypes.pyx:274-350), SURVEY.md 8f row 1.

Same attributes (`map`: (n, 5) float64 rows (mid1, mid2, contactCount, p, q); `regions`; `resolution`) and methods
(`decimate`, `contacts`, `to_matrix`) as the reference class; `decimate` - the sort / reduce-by-key - runs on the GPU
(K7, `bbk_decimate` in include/bbk.h).  The reference constructor builds its path from a hard-coded NFS template
(datatypes.pyx:26, :310); here the file name is given directly (`FithicContactMap(path, resolution)`), or use
`FithicContactMap.from_arrays(map, resolution)`.
"""
import numpy as np

from .utils import Q_LOWER_BOUND


class FithicContactMap(object):
    def __init__(self, filename, resolution=1000, chromosome=None, celltype=None):
        import pandas
        self.resolution = resolution
        self.filename = filename
        self.chromosome = chromosome
        self.celltype = celltype
        # datatypes.pyx:314: columns fragmentMid1, fragmentMid2, contactCount, p-value, q-value
        self.map = pandas.read_csv(self.filename, sep="\t", usecols=[1, 3, 4, 5, 6], engine='c', dtype='float64').values
        self.regions = np.union1d(self.map[:, 0], self.map[:, 1])

    @classmethod
    def from_arrays(cls, map5, resolution=1000, chromosome=None, celltype=None):
        self = cls.__new__(cls)
        self.resolution = resolution
        self.filename = None
        self.chromosome = chromosome
        self.celltype = celltype
        self.map = np.ascontiguousarray(map5, dtype=np.float64).reshape(-1, 5)
        self.regions = np.union1d(self.map[:, 0], self.map[:, 1])
        return self

    def decimate(self, resolution=5000):
        """datatypes.pyx:317-339: round both midpoints to the coarser grid ((int(mid) + r) / r * r - r/2, Python-2 integer
        division) and fold the rows that coincide - sum of counts, product of p, min of q, in file order."""
        import torch
        from . import _lib
        lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.BbkError("FithicContactMap.decimate needs a CUDA device (blueberry_b200 has no CPU fallback)")
        dev = torch.device("cuda:%d" % torch.cuda.current_device())
        n = int(self.map.shape[0])
        self.resolution = resolution
        if n == 0:
            return
        rows = torch.from_numpy(np.ascontiguousarray(self.map, dtype=np.float64)).to(dev)    # (n, 5) row-major, as the reference holds it
        out = torch.empty((n, 5), dtype=torch.float64, device=dev)
        n_out = torch.zeros(1, dtype=torch.int64, device=dev)
        ws = torch.empty(int(lib.bbk_decimate_workspace_bytes(n)), dtype=torch.uint8, device=dev)
        _lib.check(lib.bbk_decimate(_lib.ptr(rows), n, int(resolution), _lib.ptr(out), _lib.ptr(n_out), _lib.ptr(ws), ws.numel(),
                                    _lib.stream_ptr()), "bbk_decimate")
        g = int(n_out.item())
        if g < 0:
            raise ValueError("decimate: a midpoint is outside [0, 2^31)")
        self.map = out[:g].cpu().numpy()
        self.regions = np.union1d(self.map[:, 0], self.map[:, 1])

    def contacts(self):
        """datatypes.pyx:341-350: all contacts with a q-value <= Q_LOWER_BOUND, as (mid1, mid2) rows."""
        return self.map[self.map[:, 4] <= Q_LOWER_BOUND, :2]

    def to_matrix(self, statistic='count', n_bins=None):
        """datatypes.pyx:352-388: the chosen column as a dense 2-d matrix indexed by bin.  The reference sizes the matrix
        from a KR-norm vector it loads from a hard-coded path; here `n_bins` (default: enough for the largest midpoint)."""
        col = {'count': 2, 'p': 3, 'q': 4}.get(statistic)
        if col is None:
            raise ValueError
        r = self.resolution
        h = int(r) // 2                               # the reference's Python-2 `resolution / 2` floors
        i = ((self.map[:, 0] - h) / r).astype(np.int64)
        j = ((self.map[:, 1] - h) / r).astype(np.int64)
        d = int(n_bins) if n_bins is not None else (int(max(i.max(), j.max())) + 1 if len(i) else 0)
        matrix = np.zeros((d, d))
        matrix[i, j] = self.map[:, col]          # later rows win, like the reference's loop
        return matrix


class ContactMap(object):
    """The raw contact map next to the pass (blueberry/datatypes.pyx:31-171) as BAND RECORDS on the GPU instead of the
    reference's dense (n_bins+1)^2 matrix (impossible beyond ~25 kb on chr1): `bin1 <= bin2`, `value`.

    Same constructor inputs as the reference reads (RAWobserved rows pos1, pos2, count; the KRnorm and KRexpected
    vectors), given as paths or arrays instead of its hard-coded NFS templates; `normalize()` is the reference's
    KR balancing + observed/expected + nan_to_num in one elementwise kernel (K8); `to_dense()` rebuilds the
    reference's matrix for small maps.
    """

    def __init__(self, raw, kr_norm, kr_expected, resolution=1000, chromosome=None, celltype=None):
        import pandas
        import torch
        from . import _lib
        self.resolution = int(resolution)
        self.chromosome = chromosome
        self.celltype = celltype
        self.filename = raw if isinstance(raw, str) else None
        self.KRnorm = np.loadtxt(kr_norm) if isinstance(kr_norm, str) else np.asarray(kr_norm, dtype=np.float64)
        self.KRexpected = np.loadtxt(kr_expected) if isinstance(kr_expected, str) else np.asarray(kr_expected, dtype=np.float64)
        self.n_bins = int(self.KRnorm.shape[0])                                           # datatypes.pyx:96
        if isinstance(raw, str):
            data = pandas.read_csv(raw, delimiter="\t", engine='c', dtype='float64', header=None).values   # :101
        else:
            data = np.stack([np.asarray(c, dtype=np.float64) for c in raw], axis=1)
        lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.BbkError("ContactMap needs a CUDA device (blueberry_b200 has no CPU fallback)")
        self._dev = torch.device("cuda:%d" % torch.cuda.current_device())
        n = int(data.shape[0])
        cols = torch.from_numpy(np.ascontiguousarray(data[:, :3].T)).to(self._dev)
        self.bin1 = torch.empty(n, dtype=torch.int32, device=self._dev)
        self.bin2 = torch.empty(n, dtype=torch.int32, device=self._dev)
        self.value = torch.empty(n, dtype=torch.float64, device=self._dev)
        bad = torch.zeros(1, dtype=torch.int32, device=self._dev)
        _lib.check(lib.bbk_contact_band_ingest(_lib.ptr(cols[0]), _lib.ptr(cols[1]), _lib.ptr(cols[2]), n, self.resolution, self.n_bins,
                                               _lib.ptr(self.bin1), _lib.ptr(self.bin2), _lib.ptr(self.value), _lib.ptr(bad),
                                               _lib.stream_ptr()), "bbk_contact_band_ingest")
        if int(bad.item()):
            raise IndexError("a contact falls outside the (n_bins+1)^2 map the KRnorm vector defines")
        clean = np.nan_to_num(data)
        self.regions = np.union1d(clean[:, 0], clean[:, 1])                               # :119-120
        self.regions.sort()

    def normalize(self):
        """datatypes.pyx:143-171: value /= KRnorm[i] * KRnorm[j] * KRexpected[j - i], then nan_to_num."""
        import torch
        from . import _lib
        lib = _lib.load()
        kr = torch.from_numpy(np.ascontiguousarray(self.KRnorm, dtype=np.float64)).to(self._dev)
        ke = torch.from_numpy(np.ascontiguousarray(self.KRexpected, dtype=np.float64)).to(self._dev)
        if ke.numel() < self.n_bins:
            raise IndexError("KRexpected is shorter than KRnorm")
        bad = torch.zeros(1, dtype=torch.int32, device=self._dev)
        _lib.check(lib.bbk_contact_band_normalize(_lib.ptr(self.bin1), _lib.ptr(self.bin2), _lib.ptr(self.value), int(self.value.numel()),
                                                  _lib.ptr(kr), _lib.ptr(ke), self.n_bins, _lib.ptr(self.value), _lib.ptr(bad),
                                                  _lib.stream_ptr()), "bbk_contact_band_normalize")
        if int(bad.item()):
            raise ZeroDivisionError("float division")        # what the reference's checked division raises (datatypes.pyx:168)

    def to_dense(self):
        """The reference's `matrix` attribute: (n_bins+1)^2, both triangles; a repeated cell keeps the last record."""
        d = self.n_bins + 1
        i, j, v = self.bin1.cpu().numpy(), self.bin2.cpu().numpy(), self.value.cpu().numpy()
        matrix = np.zeros((d, d), dtype=np.float64)
        matrix[i, j] = v
        matrix[j, i] = v
        return matrix
