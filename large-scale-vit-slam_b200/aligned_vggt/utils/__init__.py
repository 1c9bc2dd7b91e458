from aligned_vggt import fall_through as _fall_through

_fall_through(__name__, __path__)  # modules not provided here resolve to the reference checkout further down sys.path
