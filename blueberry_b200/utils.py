"""Constants of blueberry/utils.py:23-26 (the band used by extract_contacts / count_band_regions)."""
Q_LOWER_BOUND = 0.01
Q_UPPER_BOUND = 0.50
HIGH_FITHIC_CUTOFF = 10000000
LOW_FITHIC_CUTOFF = 25000
