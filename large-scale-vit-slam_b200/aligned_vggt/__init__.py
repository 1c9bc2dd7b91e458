"""Drop-in `aligned_vggt` package: the modules on the hot path (models, heads.alignment_head, utils.{alignment,data,geometry})
are provided here and run in liblsvs_b200.so; every other module of the reference's package (utils.visualization, layers.*,
anything added later) keeps resolving to the reference's own file.

The reference ships `aligned_vggt/` without `__init__.py` files (a namespace package), so a regular package placed ahead of it on
`sys.path` would hide those modules.  `fall_through()` appends the same-named directories found further down `sys.path`
to a package's `__path__`: Python then looks here first and in the reference checkout second.
"""
import os
import sys


def fall_through(package_name: str, package_path: list) -> None:
    """Append every other `<entry>/<package_name as a path>` directory on sys.path to `package_path` (in sys.path order)."""
    rel = os.path.join(*package_name.split("."))
    have = {os.path.realpath(p) for p in package_path}
    for entry in list(sys.path):
        cand = os.path.join(entry or os.getcwd(), rel)
        if os.path.isdir(cand) and os.path.realpath(cand) not in have:
            package_path.append(cand)
            have.add(os.path.realpath(cand))


fall_through(__name__, __path__)
