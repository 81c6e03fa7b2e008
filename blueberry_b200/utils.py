"""Constants of blueberry/utils.py:23-26 (the band used by extract_contacts / count_band_regions)."""
Q_LOWER_BOUND = 0.01
Q_UPPER_BOUND = 0.50
HIGH_FITHIC_CUTOFF = 10000000
LOW_FITHIC_CUTOFF = 25000


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
    contact = np.array(contact_map, dtype=np.float64, copy=True)
    if contact.ndim != 2 or contact.shape[1] != 5:
        raise ValueError("contact_map must have the 5 columns mid1, mid2, contactCount, p, q")
    if alpha is not None:
        contact = contact[contact[:, 3] <= alpha]
    contact[:, 1:] = contact[:, :-1].copy()
    contact[:, 0] = chromosome
    distances = contact[:, 2] - contact[:, 1]
    contact = contact[(distances <= HIGH_FITHIC_CUTOFF) & (distances >= LOW_FITHIC_CUTOFF)]
    if regions is not None:
        from .blueberry import count_band_regions
        return contact, count_band_regions(np.ascontiguousarray(regions, dtype=np.float64))
    return contact
