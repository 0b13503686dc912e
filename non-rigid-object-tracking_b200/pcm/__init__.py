"""Host-side support code of the B200-native PC masker (ctypes binding to
libpcm_b200.so, forest/PCA export, label/bbox providers, synthetic sequences,
multi-GPU sweep).  The reference-facing plugin API lives in ../maskers/."""

import os as _os

# The sweep runs 8 - 16 sequence threads per process, each on its own CUDA stream.  With the default of 8 hardware work
# queues several streams share a queue and wait for each other's launches; 32 was measured at +10-20 % sequences/s on the
# 256-sequence sweep (profiles/README.md, round 2).  Read by the CUDA driver when it initialises, so it is set on import.
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
