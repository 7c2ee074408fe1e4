"""Types shared by the host mirror and the device binding (treeform/pixie src/pixie/common.nim)."""
from __future__ import annotations


class PixieError(Exception):
    """common.nim:4 — raised where the reference raises PixieError."""


# BlendMode ordinals, common.nim:6-29 (the C ABI passes ord(BlendMode) as int).
NormalBlend = 0
DarkenBlend = 1
MultiplyBlend = 2
ColorBurnBlend = 3
LightenBlend = 4
ScreenBlend = 5
ColorDodgeBlend = 6
OverlayBlend = 7
SoftLightBlend = 8
HardLightBlend = 9
DifferenceBlend = 10
ExclusionBlend = 11
HueBlend = 12
SaturationBlend = 13
ColorBlend = 14
LuminosityBlend = 15
MaskBlend = 16
OverwriteBlend = 17
SubtractMaskBlend = 18
ExcludeMaskBlend = 19

BLEND_MODE_NAMES = [
    "NormalBlend", "DarkenBlend", "MultiplyBlend", "ColorBurnBlend", "LightenBlend", "ScreenBlend",
    "ColorDodgeBlend", "OverlayBlend", "SoftLightBlend", "HardLightBlend", "DifferenceBlend",
    "ExclusionBlend", "HueBlend", "SaturationBlend", "ColorBlend", "LuminosityBlend", "MaskBlend",
    "OverwriteBlend", "SubtractMaskBlend", "ExcludeMaskBlend",
]


def rgbx(r, g, b, a) -> int:
    """Pack premultiplied ColorRGBX bytes the way the C ABI takes them (little-endian r,g,b,a)."""
    return (int(r) & 255) | ((int(g) & 255) << 8) | ((int(b) & 255) << 16) | ((int(a) & 255) << 24)


def rgba_to_rgbx(r, g, b, a) -> int:
    """chroma rgbx(ColorRGBA): (c*a + 127) div 255 (pinned by tests/test_images.nim:204-228)."""
    if a == 255:
        return rgbx(r, g, b, a)
    return rgbx((r * a + 127) // 255, (g * a + 127) // 255, (b * a + 127) // 255, a)


def parseHtmlColor(s: str):
    """Subset of chroma parseHtmlColor used on this path: #rgb, #rrggbb, #rrggbbaa and a few names.
    Returns straight-alpha (r, g, b, a) bytes."""
    names = {"black": (0, 0, 0, 255), "white": (255, 255, 255, 255), "red": (255, 0, 0, 255),
             "green": (0, 128, 0, 255), "blue": (0, 0, 255, 255), "none": (0, 0, 0, 0)}
    s = s.strip()
    if s.lower() in names:
        return names[s.lower()]
    if not s.startswith("#"):
        raise PixieError(f"Unsupported color {s!r}")
    h = s[1:]
    if len(h) == 3:
        h = "".join(ch * 2 for ch in h)
    if len(h) == 6:
        h += "ff"
    if len(h) != 8:
        raise PixieError(f"Invalid color {s!r}")
    return tuple(int(h[i:i + 2], 16) for i in (0, 2, 4, 6))
