"""Drop-in replacements of the reference's compiled extension modules (`model/lib/libsurf*.so`,
`model/lib/librf*.so`, installed there by src/SWD/CMakeLists.txt:6-9 and src/RF/CMakeLists.txt:6-9)."""
from . import librf, libsurf  # noqa: F401
