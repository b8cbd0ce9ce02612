#!/usr/bin/env python
"""Times every stage of the hot path on the GPU (through the host-buffer C ABI, wall clock around synchronous calls,
best of N) next to the CPU oracle on the same input, for the configurations of BASELINE.json.  Writes one JSON line
per stage.  usage: python tools/bench_stages.py [--quick] [--no-cpu]"""
import importlib, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
spx = importlib.import_module("superpixels-segmentation-gui-opencv_b200")
synth = importlib.import_module("superpixels-segmentation-gui-opencv_b200.synth")
from oracle import oracle as O

quick = "--quick" in sys.argv
cpu = "--no-cpu" not in sys.argv
O.set_threads(os.cpu_count() or 1)
PEAK = 6530.3
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def best(f, n=5):
    t = 1e30
    for _ in range(n):
        t0 = time.perf_counter(); r = f(); t = min(t, time.perf_counter() - t0)
    return t, r


def line(stage, cfg, px, bytes_per_px, t_gpu, t_cpu, extra=None):
    out = {"stage": stage, "config": cfg, "gpu_ms": t_gpu * 1e3, "gpu_mpx_s": px / 1e6 / t_gpu,
           "algorithmic_GBs": px * bytes_per_px / t_gpu / 1e9, "frac_of_hbm_peak": px * bytes_per_px / t_gpu / 1e9 / PEAK,
           "cpu_oracle_ms": None if t_cpu is None else t_cpu * 1e3, "cpu_threads": O.lib().spxo_get_threads(),
           "note": "wall clock around synchronous host-buffer C-ABI calls (includes H2D/D2H of that call)"}
    if extra:
        out.update(extra)
    print(json.dumps(out), flush=True)


def run():
    W, H = (1920, 1080) if quick else (3840, 2160)
    bgr, gt = synth.synth(W, H, 0)
    px = W * H
    # Lab conversion
    t, lab = best(lambda: spx.cvtColor(bgr, spx.COLOR_BGR2Lab))
    tc = best(lambda: O.bgr2lab8(bgr), 2)[0] if cpu else None
    line("cvtColor BGR2Lab", "%dx%d" % (W, H), px, 6, t, tc)
    # SLIC family: create + iterate(10) + sync
    for alg, name, S in ((spx.SLIC, "SLIC", 20), (spx.SLICO, "SLICO", 25), (spx.MSLIC, "MSLIC", 30)):
        def g():
            s = spx.createSuperpixelSLIC(lab, alg, S, 10.0); s.iterate(10); s.getNumberOfSuperpixels(); return s
        g()
        t, s = best(g, 3)
        s2 = spx.createSuperpixelSLIC(lab, alg, S, 10.0)
        t_it = best(lambda: (s2.reset(lab), spx.lib().spx_slic_synchronize(s2._h), 0)[2], 1)[0]
        def it_only():
            s2.reset(lab); spx.lib().spx_slic_synchronize(s2._h)
            t0 = time.perf_counter(); s2.iterate(10); spx.lib().spx_slic_synchronize(s2._h); return time.perf_counter() - t0
        t_iter = min(it_only() for _ in range(5))
        tc = None
        if cpu:
            def c():
                o = O.SuperpixelSLIC(lab, alg, S, 10.0); o.iterate(10); return o
            tc = best(c, 1)[0]
        line("%s create+iterate(10)" % name, "%dx%d S=%d" % (W, H, S), px, 70, t, tc, {"iterate10_only_ms": t_iter * 1e3,
             "iterate10_frac_of_hbm_peak": px * 70 / t_iter / 1e9 / PEAK})
        if alg == spx.SLIC:
            # connectivity + contour on the SLIC result
            L0 = s.getLabels(); k0 = s.getNumberOfSuperpixels()
            def conn():
                s.setLabels(L0, k0); t0 = time.perf_counter(); s.enforceLabelConnectivity(25); s.getNumberOfSuperpixels(); return time.perf_counter() - t0
            conn(); t = min(conn() for _ in range(5))
            tc = best(lambda: O.enforce_label_connectivity(L0, k0, 25), 1)[0] if cpu else None
            line("enforceLabelConnectivity(25)", "%dx%d, labels resident" % (W, H), px, 8, t, tc)
            L1 = s.getLabels()
            for thick in (False, True):
                t, _ = best(lambda: s.getLabelContourMask(thick), 3)
                tc = best(lambda: O.label_contour_mask(L1, thick), 1)[0] if cpu else None
                line("getLabelContourMask(thick_line=%s)" % thick, "%dx%d" % (W, H), px, 5, t, tc)
    # LSC at 8K (or 4K when quick)
    W2, H2 = (3840, 2160) if quick else (7680, 4320)
    bgr2, _ = synth.synth(W2, H2, 0)
    lab2 = spx.cvtColor(bgr2, spx.COLOR_BGR2Lab)
    def gl():
        l = spx.createSuperpixelLSC(lab2, 30, 0.075); l.iterate(10); l.getNumberOfSuperpixels(); return l
    gl(); t, l = best(gl, 2)
    tc = None
    if cpu:
        def cl():
            o = O.SuperpixelLSC(lab2, 30, 0.075); o.iterate(10); return o
        tc = best(cl, 1)[0]
    line("LSC create+iterate(10)", "%dx%d S=30 ratio=0.075" % (W2, H2), W2 * H2, 70, t, tc)
    Ll, kl = l.getLabels(), l.getNumberOfSuperpixels()
    l.enforceLabelConnectivity(25)                       # first call: the device memory pool grows (hundreds of MB)
    l.setLabels(Ll, kl)
    t0 = time.perf_counter(); l.enforceLabelConnectivity(25); k = l.getNumberOfSuperpixels(); t = time.perf_counter() - t0
    line("LSC enforceLabelConnectivity(25)", "%dx%d" % (W2, H2), W2 * H2, 8, t, None, {"K_after": k})
    if not quick:
        def gm():
            m = spx.createSuperpixelSLIC(lab2, spx.MSLIC, 30, 10.0); m.iterate(10); m.getNumberOfSuperpixels(); return m
        gm(); t, m = best(gm, 2)
        line("MSLIC create+iterate(10)", "%dx%d S=30" % (W2, H2), W2 * H2, 70, t, None)
    # SEEDS 1080p
    bgr3, _ = synth.synth(1920, 1080, 0)
    lab3 = spx.cvtColor(bgr3, spx.COLOR_BGR2Lab)
    sd = spx.createSuperpixelSEEDS(1920, 1080, 3, 2000, 4, 2, 5, False)
    sd.iterate(lab3, 10)
    t, _ = best(lambda: sd.iterate(lab3, 10), 3)
    tc = None
    if cpu:
        so = O.SuperpixelSEEDS(1920, 1080, 3, 2000, 4, 2, 5, False)
        O.set_threads(1)
        tc = best(lambda: so.iterate(lab3, 10), 1)[0]
    line("SEEDS iterate(img, 10)", "1920x1080 2000 superpixels 4 levels 5 bins", 1920 * 1080, 3 + 20 * 10, t, tc, {"cpu_threads": 1, "launches": sd.launchCount()})

    # ---- the GUI's passes over the segmentation working set, next to the cv2 calls the
    # reference makes (cv2 = OpenCV 4.13 on the host cores, its own thread pool)
    import cv2
    W, H = 3840, 2160
    bgr, _ = synth.synth(W, H, 0)
    lab = spx.cvtColor(bgr, spx.COLOR_BGR2Lab)
    s = spx.createSuperpixelSLIC(lab, spx.SLIC, 20, 10.0); s.iterate(10); s.enforceLabelConnectivity(25)
    labels = s.getLabels(); contour = s.getLabelContourMask(False)
    ses = spx.Session(W, H)
    t, _ = best(lambda: ses.begin(labels, contour, (0, 0, 255)), 3)
    ses.upload(0, bgr)
    line("session begin (labels + grid upload, cell index)", "3840x2160", W * H, 4 + 1 + 3 + 12, t, None)
    rng = np.random.default_rng(0)
    pts = [(int(rng.integers(0, W)), int(rng.integers(0, H))) for _ in range(64)]
    t0 = time.perf_counter()
    for (x, y) in pts:
        ses.whichCell(x, y, True, (0, 255, 0), 1)
    t = (time.perf_counter() - t0) / len(pts)
    mask = np.zeros((H, W, 3), np.uint8); lm = np.zeros((H, W), np.int32)

    def cv_click(x, y):
        m = cv2.compare(labels, int(labels[y, x]), cv2.CMP_EQ)       # Mat1b superpixel_mask = labels == label
        mask[m != 0] = 0; lm[m != 0] = 0                               # setTo(0, superpixel_mask) x2
        mask[m != 0] = (0, 255, 0); lm[m != 0] = 1                     # setTo(color / label, superpixel_mask)
    tc = None
    if cpu:
        t0 = time.perf_counter()
        for (x, y) in pts[:8]:
            cv_click(x, y)
        tc = (time.perf_counter() - t0) / 8
    line("WhichCell click (one cell painted)", "3840x2160, ~400 px cells", W * H, 4, t, tc, {"cpu": "cv2.compare + numpy masked stores", "cpu_threads": cv2.getNumThreads()})
    for (vw, vh) in ((1920, 1080), (3840, 2160)):
        vp = (0, 0, vw, vh)
        t, _ = best(lambda: ses.compose(vp, image=50, mask=50, grid=100, selection=1, holes=True), 5)
        tc = None
        if cpu:
            img_c = np.ascontiguousarray(bgr[:vh, :vw]); mk = ses.download(1); gr = ses.download(2); sel = ses.download(3)

            def cv_show():
                d = np.zeros((vh, vw, 3), np.uint8)
                d = cv2.addWeighted(d, 0.5, img_c, 0.5, 0)
                d = cv2.addWeighted(d, 1, np.ascontiguousarray(mk[:vh, :vw]), 0.5, 0)
                d = cv2.addWeighted(d, 1, np.ascontiguousarray(gr[:vh, :vw]), 1.0, 0)
                sd = cv2.dilate(np.ascontiguousarray(sel[:vh, :vw]), cv2.getStructuringElement(cv2.MORPH_RECT, (3, 3)))
                cv2.copyTo(sd, sd, d)
                hm = cv2.inRange(mk, (0, 0, 0), (0, 0, 0))
                holes = np.zeros((H, W, 3), np.uint8); holes[hm != 0] = (255, 255, 255)
                d = cv2.addWeighted(d, 1, np.ascontiguousarray(holes[:vh, :vw]), 1, 0)
                return d, cv2.countNonZero(hm)
            tc = best(cv_show, 3)[0]
        line("ShowSegmentation (image+mask+grid+selection+holes)", "viewport %dx%d of 3840x2160" % (vw, vh), vw * vh, 15, t, tc,
             {"cpu": "cv2 4.13 call sequence of mainwindow.cpp:1936-1969", "cpu_threads": cv2.getNumThreads()})
    ses.release()
    # ---- pre-processing that feeds the path next to the cv2 calls of on_button_apply_clicked
    t, _ = best(lambda: spx.preprocess(bgr, equalize=True, normalize=True, color_balance=2, gaussian_blur=True), 5)
    tc = None
    if cpu:
        def cv_pre():
            a = cv2.cvtColor(bgr, cv2.COLOR_BGR2YCrCb); ch = list(cv2.split(a)); ch[0] = cv2.equalizeHist(ch[0])
            a = cv2.cvtColor(cv2.merge(ch), cv2.COLOR_YCrCb2BGR)
            a = cv2.normalize(a, None, 1, 255, cv2.NORM_MINMAX, -1)
            ch = list(cv2.split(a))
            for i in range(3):
                flat = np.sort(ch[i].reshape(1, -1), axis=1); n = flat.shape[1]
                lo, hi = int(flat[0, int(n * 0.01)]), int(flat[0, min(int(np.ceil(n * 0.99)), n - 1)])
                c = np.clip(ch[i], lo, hi); ch[i] = cv2.normalize(c, None, 0, 255, cv2.NORM_MINMAX)
            return cv2.GaussianBlur(cv2.merge(ch), (3, 3), 0, 0)
        tc = best(cv_pre, 2)[0]
    line("preprocess (equalize+normalize+balance 2%+blur 3x3)", "3840x2160", W * H, 6 * 4 + 3 * 3, t, tc,
         {"cpu": "cv2 4.13 call sequence of mainwindow.cpp:2028-2031 (sort by numpy)", "cpu_threads": cv2.getNumThreads()})


run()
