"""The gather in front of the genome-wide q-value step: blueberry/utils.py:23-90 on the device.

    extract_contacts(celltype, chromosome, resolution, alpha=None, n_regions=None)      utils.py:31-90, same signature
    extract_contacts_from_map(contact_map, chromosome, alpha=None, regions=None)        the same on an in-memory table
    genome_qvalues(maps, alpha=None)     the driver the authors scripted around these pieces:
                                         extract every chromosome -> count_band_regions -> Benjamini-Hochberg over all
                                         chromosomes with n = sum of the band counts (blueberry.pyx:40-91)

The filters run as one order-preserving stream compaction on the GPU (bbk_extract_contacts), count_band_regions is K6,
the q-values K5; there is no CPU fallback.
"""
import os

Q_LOWER_BOUND = 0.01
Q_UPPER_BOUND = 0.50
HIGH_FITHIC_CUTOFF = 10000000
LOW_FITHIC_CUTOFF = 25000

# Where extract_contacts finds a chromosome's result file: format(celltype, chromosome, resolution), like the module global
# of datatypes.pyx:26 (a lab path there; here relative to $BLUEBERRY_DATA or the working directory - assign to change it).
DATA_DIR = os.path.join(os.environ.get("BLUEBERRY_DATA", "."), "{2}", "{0}.chr{1}.spline_pass1.res{2}.significances.txt.gz")


def _extract_on_device(dmap, n, chromosome, alpha, dev):
    """(n, 5) device table -> (kept rows as a device tensor (k, 5), k)."""
    import torch
    from . import _lib
    lib = _lib.load()
    out = torch.empty((max(n, 1), 5), dtype=torch.float64, device=dev)
    n_out = torch.zeros(1, dtype=torch.int64, device=dev)
    ws = torch.empty(int(lib.bbk_extract_workspace_bytes(n)), dtype=torch.uint8, device=dev)
    _lib.check(lib.bbk_extract_contacts(_lib.ptr(dmap), n, float(chromosome), float(alpha) if alpha is not None else 0.0,
                                        0 if alpha is None else 1, float(LOW_FITHIC_CUTOFF), float(HIGH_FITHIC_CUTOFF), _lib.ptr(out), n,
                                        _lib.ptr(n_out), _lib.ptr(ws), ws.numel(), _lib.stream_ptr()), "bbk_extract_contacts")
    k = int(n_out.item())
    return out[:k], k


def _device():
    import torch
    from . import _lib
    if not torch.cuda.is_available():
        raise _lib.BbkError("blueberry_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def extract_contacts_from_map(contact_map, chromosome, alpha=None, regions=None):
    """utils.extract_contacts (utils.py:31-90) on an in-memory Fit-Hi-C result table.

    The reference loads `FithicContactMap(celltype, chromosome, resolution).map` from a lab path
    (datatypes.pyx:308-315); here the (n, 5) float64 table - columns mid1, mid2, contactCount, p, q, as in
    the significances file - is passed in.  Steps as in the reference: keep p <= alpha (:72-73), shift the
    columns right and put the chromosome first (:76-77), keep LOW_FITHIC_CUTOFF <= mid2 - mid1 <=
    HIGH_FITHIC_CUTOFF (:80-83).  With `regions` (float64 midpoints) also returns
    count_band_regions(regions) computed on the device (:87-88).
    """
    import numpy as np
    import torch
    contact = np.ascontiguousarray(contact_map, dtype=np.float64)
    if contact.ndim != 2 or contact.shape[1] != 5:
        raise ValueError("contact_map must have the 5 columns mid1, mid2, contactCount, p, q")
    dev = _device()
    n = int(contact.shape[0])
    kept, _ = _extract_on_device(torch.from_numpy(contact).to(dev), n, chromosome, alpha, dev)
    contact = kept.cpu().numpy()
    if regions is not None:
        from .blueberry import count_band_regions
        return contact, count_band_regions(np.ascontiguousarray(regions, dtype=np.float64))
    return contact


def extract_contacts(celltype, chromosome, resolution, alpha=None, n_regions=None):
    """Extract contacts from a given chromosome, and number of regions in the band (utils.py:31-90, same signature).

    Returns the (k, 5) rows (chromosome, mid1, mid2, contactCount, p) with p <= alpha inside the 25 kb - 10 Mb band and,
    with n_regions, count_band_regions over the chromosome's regions.  Like the reference, a chromosome whose file cannot
    be read gives (numpy.zeros((0, 5)), 0) and a message instead of an exception (:65-67)."""
    import numpy as np
    from .datatypes import FithicContactMap
    print("CPU [{}]: Extracting {} chr{}".format(chromosome, celltype, chromosome))
    try:
        contact_map = FithicContactMap(DATA_DIR.format(celltype, chromosome, resolution), resolution, chromosome, celltype)
    except Exception as e:                                     # utils.py:65-67
        print("CPU [{}]: {}".format(chromosome, e))
        return np.zeros((0, 5)), 0
    if n_regions:
        return extract_contacts_from_map(contact_map.map, chromosome, alpha, regions=contact_map.regions)
    return extract_contacts_from_map(contact_map.map, chromosome, alpha)


def genome_qvalues(maps, alpha=None):
    """Genome-wide q-values the way the authors composed them from these pieces: for every chromosome the
    contacts with p <= alpha inside the band (extract_contacts) and the number of region pairs inside the band
    (count_band_regions over union1d(mid1, mid2)), then Benjamini-Hochberg over the p-values of all chromosomes with
    n = the sum of the band counts (blueberry.benjamini_hochberg on the sorted p, blueberry.pyx:40-75).

    maps: {chromosome: (n, 5) table or FithicContactMap}.  Everything runs on the device: one stream compaction per
    chromosome, K6 per chromosome, one K5 launch over the concatenated p (ranked on the device, so nothing is sorted on
    the host).  Returns (contacts, q, n): contacts (m, 5) rows (chromosome, mid1, mid2, contactCount, p) chromosome by
    chromosome in the order given, q (m,) their q-values, n the total band count."""
    import numpy as np
    import torch
    from . import _lib
    from .blueberry import count_band_regions
    lib = _lib.load()
    dev = _device()
    parts, n_total = [], 0
    for chrom, m in maps.items():
        table = np.ascontiguousarray(getattr(m, "map", m), dtype=np.float64)
        if table.ndim != 2 or table.shape[1] != 5:
            raise ValueError("every map must have the 5 columns mid1, mid2, contactCount, p, q")
        regions = getattr(m, "regions", None)
        if regions is None:
            regions = np.union1d(table[:, 0], table[:, 1])                       # datatypes.pyx:315
        n_total += count_band_regions(np.ascontiguousarray(regions, dtype=np.float64))
        kept, _ = _extract_on_device(torch.from_numpy(table).to(dev), int(table.shape[0]), chrom, alpha, dev)
        parts.append(kept)
    if not parts:
        return np.zeros((0, 5)), np.zeros(0), 0
    contacts = torch.cat(parts) if len(parts) > 1 else parts[0]
    m = int(contacts.shape[0])
    if m == 0:
        return np.zeros((0, 5)), np.zeros(0), n_total
    p = torch.empty((m + 1) & ~1, dtype=torch.float64, device=dev)[:m].copy_(contacts[:, 4])
    q = torch.empty((m + 1) & ~1, dtype=torch.float64, device=dev)[:m]
    ws = torch.empty(int(lib.bbk_bh_workspace_bytes(m)), dtype=torch.uint8, device=dev)
    _lib.check(lib.bbk_bh_qvalues(_lib.ptr(p), m, int(n_total), _lib.BH_UNSORTED, None, _lib.ptr(q), None, _lib.ptr(ws), ws.numel(),
                                  _lib.stream_ptr()), "bbk_bh_qvalues")
    return contacts.cpu().numpy(), q.cpu().numpy(), n_total
