"""Copies the reference's own golden fixtures for the raster hot path into tests/golden/.

Run in the build container (where /root/reference exists).  The copied files are test FIXTURES
(golden PNGs the reference's tests compare against, SURVEY.md section 4, and the tiger SVG input of
BASELINE config 2), not reference source code.  /root/reference does not exist on the GPU box,
so tests and bench.py read only the copies committed here.
"""
import os
import shutil
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

PATHS = """pathStroke1 pathStroke2 pathStroke3 pathBlackRectangle pathBlackRectangleZ pathYellowRectangle
pathRedRectangle pathBottomArc pathHeart pathRotatedArc pathInvertedCornerArc pathCornerArc pixelScale
boxRound boxBevel boxMiter ButtCap RoundCap SquareCap dashes selfclosing rectExcludeMask rectExcludeMaskAA
rectMask rectMaskAA rectMaskStroke opacityFill opacityStroke pathStroke1Big path1pxCover path0pxCover
polygon3 polygon4 polygon5 polygon6 polygon7 polygon8 pathSwish miterLimit_10deg_2.00num
miterLimit_145deg_2.00num miterLimit_155deg_2.00num miterLimit_165deg_2.00num miterLimit_165deg_10.00num
miterLimit_145deg_3.32num miterLimit_145deg_3.33num""".split()

FILES = [(f"tests/paths/{n}.png", f"paths_{n}.png") for n in PATHS] + [
    ("tests/images/imageblur20.png", "images_imageblur20.png"),
    ("tests/images/imageblur20oob.png", "images_imageblur20oob.png"),
    ("tests/contexts/blendmode_1.png", "contexts_blendmode_1.png"),
    ("examples/heart.png", "examples_heart.png"),
    ("examples/shadow.png", "examples_shadow.png"),
    ("examples/masking.png", "examples_masking.png"),
    ("examples/blur.png", "examples_blur.png"),
    ("examples/data/trees.png", "examples_data_trees.png"),
    ("examples/data/tiger.svg", "tiger.svg"),
    ("tests/fileformats/svg/masters/Ghostscript_Tiger.png", "svg_masters_Ghostscript_Tiger.png"),
] + [(f"tests/images/maskClearsOnDraw{i}.png", f"images_maskClearsOnDraw{i}.png") for i in range(5)]

# draw with any transform, minify / magnify, non-solid paints (tests/test_images_draw.nim, test_images.nim,
# test_paints.nim) and the input images those tests read
PAINTS = """paintSolid paintImage paintImageOpacity paintImageTiled paintImageTiledOpacity gradientLinear
gradientLinear2 gradientRadial gradientAngular gradientAngularOpacity fillImagePaint fillTiledImagePaint""".split()
DRAWS = """rotate0 rotate90 rotate180 rotate270 rotate360 scaleHalf fillOptimization fillOptimization2 flipped1
minifiedBy2 magnifiedBy2 minifiedBy4 magnifiedBy4 minifiedMandrill turtle turtle@10x rock""".split()
MASTERS = [f"smooth{i}" for i in range(1, 13)] + ["minify_odd", "rock_minified", "rock_minified2"]
FILES += [(f"tests/paths/{n}.png", f"paths_{n}.png") for n in PAINTS]
FILES += [(f"tests/images/{n}.png", f"images_{n}.png") for n in DRAWS]
FILES += [(f"tests/images/masters/{n}.png", f"images_masters_{n}.png") for n in MASTERS]
FILES += [("tests/fileformats/png/mandrill.png", "fileformats_png_mandrill.png")]

for src, dst in FILES:
    shutil.copyfile(os.path.join(REF, src), os.path.join(HERE, dst))
    print(dst)
