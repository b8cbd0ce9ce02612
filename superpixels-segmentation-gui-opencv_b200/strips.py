"""Row-strip SLIC: ONE image over several GPUs (BASELINE.json north-star: gigapixel images).

One process per GPU (torch.distributed / NCCL).  Rank r owns a band of rows (a multiple of 32, so bands are whole tile
rows) and uploads those rows plus two halo rows.  Cluster state is replicated; per iteration ONE exchange step sums the
integer accumulators (pixel sums per cluster) over the ranks -- only clusters near a cut receive pixels from two ranks,
the sums are integers, so the result is bit-identical to the single-GPU object.

The exchange step runs over NVLink peer memory: every rank's exchange buffer is mapped into its peers through CUDA IPC,
the library's own kernels store partial sums into the peers' buffers, publish a flag, wait for the peers' flags and add
(``spx_slic_strip_exchange``; csrc/slic.cu).
  * ``exchange="neighbour"`` (default): the cluster state is PARTITIONED by band -- a rank maintains the clusters whose seed
    row lies within 5S rows of its band and exchanges only with rank-1 / rank+1, over the few seed rows of clusters both
    maintain (``spx_slic_strip_partition``).  Exact while no centre moves more than 4S rows from its seed row; the centre
    update checks that on the device and iterate() reruns the whole object with the all-to-all exchange if it ever fails
    (an empty cluster jumping to the origin is the one way to get there).  Thin bands fall back to "peer" at construction.
  * ``exchange="peer"``: replicated cluster state, every rank stores the index range its band touched into every peer.
  * ``exchange="nccl"``: replicated state, NCCL all-reduce on the library's device pointers, for comparison.
torch is plumbing: the process group carries the 64-byte IPC handles, the one-off assembly of the seeds and the barriers.
"""
import ctypes as C

import numpy as np

from . import SLIC, SpxError, _ck, _p, lib


class _State(C.Structure):
    _fields_ = [("kx", C.c_void_p), ("ky", C.c_void_p), ("kc", C.c_void_p), ("ikx", C.c_void_p), ("iky", C.c_void_p),
                ("acc", C.c_void_p), ("K", C.c_int), ("Kcap", C.c_int), ("channels", C.c_int),
                ("labels", C.c_void_p), ("label_pitch", C.c_int)]


class _DevArray:
    """zero-copy view of a raw device pointer for torch.as_tensor"""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def row_bands(height, world, tile=32):
    """[own_row0, own_row1) per rank: whole tile rows, as even as possible; trailing ranks may get nothing for tiny images"""
    tiles = (height + tile - 1) // tile
    out = []
    for r in range(world):
        t0, t1 = tiles * r // world, tiles * (r + 1) // world
        out.append((min(t0 * tile, height), min(t1 * tile, height)))
    return out


def halo_rows(own0, own1, height, halo=2):
    """rows a rank has to upload: its own band plus `halo` rows on every inner side"""
    return max(own0 - halo, 0), min(own1 + halo, height)


class StripSLIC:
    """cv::ximgproc::SuperpixelSLIC (algorithm SLIC) over the row band of this rank.  `image` is the whole 8UC3 image
    or a callable rows(row0, row1) -> ndarray, so a rank never needs more than its band in host memory."""

    def __init__(self, image, region_size, ruler, dist, device, width=None, height=None, exchange="neighbour"):
        import torch
        if exchange not in ("neighbour", "peer", "nccl"):
            raise SpxError(-5, "StripSLIC: exchange must be 'neighbour', 'peer' or 'nccl'")
        self.exchange = exchange
        self._args = (image, region_size, ruler, dist, device, width, height)
        self._iters_done = 0
        self.torch, self.dist = torch, dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        if callable(image):
            self.w, self.h = int(width), int(height)
            rows = image
        else:
            img = np.ascontiguousarray(image, dtype=np.uint8)
            self.h, self.w = img.shape[:2]
            rows = lambda a, b: img[a:b]
        self.own0, self.own1 = row_bands(self.h, self.world)[self.rank]
        if self.own1 <= self.own0:
            raise SpxError(-5, "StripSLIC: more ranks than tile rows")
        self.row0, self.row1 = halo_rows(self.own0, self.own1, self.h)
        band = np.ascontiguousarray(rows(self.row0, self.row1))
        torch.cuda.set_device(device)
        # one explicit stream for the library's kernels AND the NCCL reductions, so that they are ordered
        # (the legacy default stream would make the library create a private stream of its own)
        self.stream = torch.cuda.Stream(device)
        self._h = C.c_void_p(None)
        L = lib()
        L.spx_slic_create_strip.argtypes = [C.c_void_p, C.c_size_t] + [C.c_int] * 5 + [C.c_float] + [C.c_int] * 4 + [C.c_void_p, C.POINTER(C.c_void_p)]
        L.spx_slic_device_state.argtypes = [C.c_void_p, C.POINTER(_State)]
        for f in (L.spx_slic_strip_rebin, L.spx_slic_strip_assign, L.spx_slic_strip_update):
            f.argtypes = [C.c_void_p]
        L.spx_slic_get_label_rows.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int]
        _ck(L.spx_slic_create_strip(_p(band), self.w * 3, self.w, self.h, 3, SLIC, int(region_size), float(ruler), self.row0, self.row1,
                                    self.own0, self.own1, C.c_void_p(self.stream.cuda_stream), C.byref(self._h)))
        st = _State()
        _ck(L.spx_slic_device_state(self._h, C.byref(st)))
        self.K = st.K
        dev = torch.device("cuda", device)
        view = lambda ptr, n, ts: torch.as_tensor(_DevArray(ptr, n, ts), device=dev)
        self._seed_f = [view(st.kx, st.Kcap, "<f4"), view(st.ky, st.Kcap, "<f4"), view(st.kc, 3 * st.Kcap, "<f4")]
        self._seed_i = [view(st.ikx, st.Kcap, "<i4"), view(st.iky, st.Kcap, "<i4")]
        self._acc = view(st.acc, 6 * st.Kcap, "<i8")
        self._lp = st.label_pitch
        self._labels = view(st.labels, self.h * st.label_pitch, "<i4")     # the device's whole-image label map, pitched rows
        self._gathered = False
        # assemble the seeds: exactly one rank contributed a non-zero entry per cluster, so the sum is exact
        with torch.cuda.stream(self.stream):
            for t in self._seed_f + self._seed_i:
                dist.all_reduce(t, op=dist.ReduceOp.SUM)
        _ck(L.spx_slic_strip_rebin(self._h))
        self.exchanged_bytes = 0
        self._core = (0, self.K)
        if self.exchange == "neighbour" and self.world > 1:
            L.spx_slic_strip_partition.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.c_int, C.c_int, C.POINTER(C.c_int)]
            flat = [v for b in row_bands(self.h, self.world) for v in b]
            own = (C.c_int * len(flat))(*flat)
            info = (C.c_int * 4)()
            rc = L.spx_slic_strip_partition(self._h, own, self.world, self.rank, info)
            ok = torch.tensor([1 if rc == 0 else 0], device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)              # (the geometry is the same on every rank, so is rc; belt and braces)
            if int(ok.item()) == 1:
                self._core = (info[2], info[3])
            else:
                if rc == 0:
                    raise SpxError(-5, "StripSLIC: the ranks disagree about the strip partition")
                self.exchange = "peer"                              # bands too thin: replicated state, all-to-all exchange
        elif self.exchange == "neighbour":
            self.exchange = "peer"
        if self.exchange in ("peer", "neighbour") and self.world > 1:
            L.spx_slic_strip_peer_alloc.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_size_t)]
            L.spx_slic_strip_peer_open.argtypes = [C.c_void_p, C.c_void_p]
            L.spx_slic_strip_exchange.argtypes = [C.c_void_p]
            mine = C.create_string_buffer(64)
            nbytes = C.c_size_t(0)
            _ck(L.spx_slic_strip_peer_alloc(self._h, self.world, self.rank, mine, C.byref(nbytes)))
            handles = [None] * self.world
            dist.all_gather_object(handles, mine.raw)              # 64 bytes per rank
            allh = C.create_string_buffer(b"".join(handles), 64 * self.world)
            _ck(L.spx_slic_strip_peer_open(self._h, allh))
            self.stream.synchronize()
            dist.barrier()                                         # every buffer is zeroed and mapped before the first push

    def iterate(self, num_iterations=10):
        L = lib()
        peer = self.exchange in ("peer", "neighbour") and self.world > 1
        self._iters_done += int(num_iterations)
        self._iterate(L, peer, num_iterations)
        bad = 0
        if self.exchange == "neighbour" and self.world > 1:
            try:
                self.exchangedBytes()                              # synchronises and reports the device-side drift check
            except SpxError as e:
                if e.code != -215:
                    raise
                bad = 1
            t = self.torch.tensor([bad], device=self._acc.device)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            bad = int(t.item())
        if bad:
            # a centre left the rows the partition covers: redo everything so far with replicated state (exact in every case)
            image, region_size, ruler, dist, device, width, height = self._args
            n = self._iters_done
            self.release()
            self.__init__(image, region_size, ruler, dist, device, width, height, exchange="peer")
            self.iterate(n)

    def _iterate(self, L, peer, num_iterations):
        if peer:
            L.spx_slic_strip_iteration.argtypes = [C.c_void_p]
            for _ in range(int(num_iterations)):
                _ck(L.spx_slic_strip_iteration(self._h))           # assign + exchange + update, one call
            return
        for _ in range(int(num_iterations)):
            _ck(L.spx_slic_strip_assign(self._h))
            if peer:
                _ck(L.spx_slic_strip_exchange(self._h))            # the one exchange step of the path: NVLink peer stores + flags
            else:
                with self.torch.cuda.stream(self.stream):
                    self.dist.all_reduce(self._acc, op=self.dist.ReduceOp.SUM)
                self.exchanged_bytes += self._acc.numel() * 8
            _ck(L.spx_slic_strip_update(self._h))

    def getNumberOfSuperpixels(self):
        return self.K

    def exchangedBytes(self):
        """bytes this rank has sent in exchange steps so far (peer modes: counted on the device by the push kernel)"""
        if self.exchange in ("peer", "neighbour") and self.world > 1:
            L = lib()
            L.spx_slic_strip_exchanged_bytes.argtypes = [C.c_void_p, C.POINTER(C.c_ulonglong)]
            v = C.c_ulonglong(0)
            _ck(L.spx_slic_strip_exchanged_bytes(self._h, C.byref(v)))
            return int(v.value)
        return self.exchanged_bytes

    def getLabelRows(self):
        """the owned rows [own0, own1) of the label map"""
        out = np.empty((self.own1 - self.own0, self.w), np.int32)
        _ck(lib().spx_slic_get_label_rows(self._h, _p(out), self.w * 4, self.own0, self.own1))
        return out

    def gatherLabels(self):
        """the bands of all ranks into rank 0's device label map, device to device (NCCL send / recv of whole pitched rows);
        collective"""
        bands = row_bands(self.h, self.world)
        with self.torch.cuda.stream(self.stream):
            if self.rank == 0:
                for r in range(1, self.world):
                    a, b = bands[r]
                    if b > a:
                        self.dist.recv(self._labels[a * self._lp:b * self._lp], src=r)
            elif self.own1 > self.own0:
                self.dist.send(self._labels[self.own0 * self._lp:self.own1 * self._lp], dst=0)
        self.stream.synchronize()
        self._gathered = True

    def enforceLabelConnectivity(self, min_element_size=25):
        """mainwindow.cpp:2108 on a strip object: connectivity is not strip-parallel, so the bands are
        gathered on rank 0's device -- no host staging -- and the single-device kernels run there; collective.  Afterwards
        rank 0 holds the label map (getLabels / getLabelContourMask return None on the other ranks)."""
        if not self._gathered:
            self.gatherLabels()
        k = [self.K]
        if self.rank == 0:
            L = lib()
            _ck(L.spx_slic_enforce_label_connectivity(self._h, int(min_element_size)))
            n = C.c_int(0)
            _ck(L.spx_slic_get_number_of_superpixels(self._h, C.byref(n)))
            k = [int(n.value)]
        self.dist.broadcast_object_list(k, src=0)
        self.K = k[0]

    def getLabelContourMask(self, thick_line=True):
        """mainwindow.cpp:2111; rank 0 returns the mask of the whole image, the others None; collective"""
        if not self._gathered:
            self.gatherLabels()
        if self.rank != 0:
            return None
        out = np.empty((self.h, self.w), np.uint8)
        _ck(lib().spx_slic_get_label_contour_mask(self._h, _p(out), self.w, 1 if thick_line else 0))
        return out

    def getLabels(self):
        """the whole label map.  Before connectivity: gathered on every rank through the host (for tests; a gigapixel caller
        keeps the bands apart).  After enforceLabelConnectivity / gatherLabels: rank 0's device map (None elsewhere)."""
        if self._gathered:
            if self.rank != 0:
                return None
            out = np.empty((self.h, self.w), np.int32)
            _ck(lib().spx_slic_get_labels(self._h, _p(out), self.w * 4))
            return out
        mine = self.getLabelRows()
        parts = [None] * self.world
        self.dist.all_gather_object(parts, mine)
        return np.concatenate(parts, 0)

    def synchronize(self):
        self.stream.synchronize()

    def getCentres(self):
        """(K, 5) centres c0 c1 c2 x y.  Neighbour mode: every rank contributes the clusters of its own band; collective."""
        t = self.torch
        self.stream.synchronize()
        K, Kcap = self.K, self._seed_f[0].numel()
        kc = self._seed_f[2].view(3, Kcap)[:, :K].t()
        mine = t.cat([kc, self._seed_f[0][:K, None], self._seed_f[1][:K, None]], 1).cpu().numpy()
        if self.exchange != "neighbour" or self.world == 1:
            return mine
        parts = [None] * self.world
        self.dist.all_gather_object(parts, (self._core, mine[self._core[0]:self._core[1]]))
        out = mine.copy()
        for (lo, hi), v in parts:
            out[lo:hi] = v
        return out

    def release(self, barrier=True):
        """collective in the peer modes: no rank may free its exchange buffer while a peer can still write to it"""
        if self._h.value:
            if barrier and self.exchange in ("peer", "neighbour") and self.world > 1:
                self.stream.synchronize()
                self.dist.barrier()
            lib().spx_slic_destroy(self._h)
            self._h = C.c_void_p(None)

    def __del__(self):
        try:
            self.release(barrier=False)
        except Exception:
            pass
