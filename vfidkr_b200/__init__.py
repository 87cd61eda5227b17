"""Import shim: `import vfidkr_b200` loads the package that lives in
`video-frame-interpolation-based-on-deformable-kernel-region_b200/` (a directory name Python cannot import)."""
from pathlib import Path as _Path

_real = _Path(__file__).resolve().parent.parent / "video-frame-interpolation-based-on-deformable-kernel-region_b200"
__path__ = [str(_real)]
__file__ = str(_real / "__init__.py")
exec(compile((_real / "__init__.py").read_text(), __file__, "exec"))
