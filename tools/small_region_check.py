import importlib, os, sys
sys.path.insert(0, '.')
import numpy as np
spx = importlib.import_module("superpixels-segmentation-gui-opencv_b200")
synth = importlib.import_module("superpixels-segmentation-gui-opencv_b200.synth")
from oracle import oracle as O
O.set_threads(os.cpu_count() or 1)
img, _ = synth.synth(3840, 2160, 4)
lab = O.bgr2lab8(img)
ok = True
for alg, name in ((spx.SLIC, "SLIC"), (spx.SLICO, "SLICO"), (spx.MSLIC, "MSLIC")):
    for S in (7, 8, 10, 13):
        a = O.SuperpixelSLIC(lab, alg, S, 10.0); a.iterate(5)
        b = spx.createSuperpixelSLIC(lab, alg, S, 10.0); b.iterate(5)
        same = bool((a.getLabels() == b.getLabels()).all() and (a.getCentres() == b.getCentres()).all())
        print(name, "S", S, "K", b.getNumberOfSuperpixels(), "identical to oracle:", same)
        ok &= same
sys.exit(0 if ok else 1)
