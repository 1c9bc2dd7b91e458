"""Host-side plumbing of the B200-native hot path: ctypes binding of the C ABI (native), the in-tree
nvcc build (build), the chunk scheduler (scheduler).  The reference-facing API lives in the sibling
``aligned_vggt`` package, which mirrors the reference's module paths."""
