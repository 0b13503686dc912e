"""Host-side support code of the B200-native PC masker (ctypes binding to
libpcm_b200.so, forest/PCA export, label/bbox providers, synthetic sequences,
multi-GPU sweep).  The reference-facing plugin API lives in ../maskers/."""
