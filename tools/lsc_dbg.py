import importlib, os, sys, time
sys.path.insert(0, "/root/repo")
spx = importlib.import_module("superpixels-segmentation-gui-opencv_b200")
synth = importlib.import_module("superpixels-segmentation-gui-opencv_b200.synth")
for (W,H) in ((1920,1080),(3840,2160),(7680,4320)):
    bgr,_ = synth.synth(W,H,0)
    lab = spx.cvtColor(bgr, spx.COLOR_BGR2Lab)
    l = spx.createSuperpixelLSC(lab, 30, 0.075); l.iterate(10)
    t0=time.time(); l.enforceLabelConnectivity(25); k=l.getNumberOfSuperpixels(); print(W,H,"enforce %.1f ms"%((time.time()-t0)*1e3), k, flush=True)
