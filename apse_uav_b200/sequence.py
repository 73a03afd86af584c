"""Native sequence post-pass: aruco_detect.py:598-782 (marker gating, marker-length recurrence, LED read-out, vehicle
distances) and the CSV rows of :146-185 for a whole sequence, on top of the apse_sequence_* entry points of
libapse_b200.so (csrc/sequence.cu).

The frame loop of the reference is sequential only through a handful of scalars (markerLength, previous marker centres,
detected flags); everything heavy in it -- the cv2.projectPoints calls for the LED strip and the vehicle outlines -- feeds
outputs only.  So the post-pass is: a native O(markers) scan on the host (microseconds per thousand frames), ONE batched
second pose launch with the exact per-frame marker lengths, the scan again, ONE launch for all projection jobs, rows.
postpass.py holds the same logic as readable per-frame Python (the mirror of the reference's text); the parity tests check
the two against each other and against the reference script's own CSV.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import ApseError, SeqConfig, SEQ_JOB_DTYPE, SEQ_RESULT_DTYPE, SEQ_ROW_DTYPE


def seq_config(start_frame=1, step_frame=1, leds=False, leds_threshold=None, width=3840, height=2160) -> SeqConfig:
    c = SeqConfig()
    _lib.load().apse_seq_config_default(C.byref(c))
    c.start_frame, c.step_frame, c.width, c.height = int(start_frame), int(step_frame), int(width), int(height)
    c.leds = 1 if leds else 0
    c.leds_threshold = -1 if leds_threshold is None else int(leds_threshold)
    return c


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def new_state():
    """State of one pass of a chunked scan (apse_seq_state: the reference's module globals, zeroed)."""
    return np.zeros(512, np.uint8)


def scan(cfg: SeqConfig, n, ids, corners, rvec, tvec, rescale_tvec=False, want_rows=True, state=None, frame0=0):
    """apse_sequence_scan on host arrays n [F], ids [F,M], corners [F,M,4,2], rvec / tvec [F,M,3].
    Returns (lengths [F] float64, rows | None, jobs | None).
    state (new_state()) + frame0: these are frames frame0 .. frame0 + F - 1 of a sequence scanned chunk by chunk
    (apse_sequence_scan_chunk); a job-list overflow is retried from a copy of the state."""
    lib = _lib.load()
    if state is not None:
        return _scan_chunk(lib, cfg, state, int(frame0), n, ids, corners, rvec, tvec, rescale_tvec, want_rows)
    n = np.ascontiguousarray(n, np.int32)
    F = int(n.shape[0])
    ids = np.ascontiguousarray(ids, np.int32).reshape(F, -1)
    M = int(ids.shape[1]) if F else 1
    corners = np.ascontiguousarray(corners, np.float32).reshape(F, M, 4, 2)
    rvec = np.ascontiguousarray(rvec, np.float64).reshape(F, M, 3)
    tvec = np.ascontiguousarray(tvec, np.float64).reshape(F, M, 3)
    lengths = np.empty(F, np.float64)
    rows = jobs = None
    nj = C.c_int(0)
    if want_rows:
        rows = np.zeros(F, SEQ_ROW_DTYPE)
        cap = 4 * F + 64
        while True:
            jobs = np.zeros(cap, SEQ_JOB_DTYPE)
            rc = lib.apse_sequence_scan(C.byref(cfg), F, M, _ptr(n), _ptr(ids), _ptr(corners), _ptr(rvec), _ptr(tvec),
                                        1 if rescale_tvec else 0, _ptr(lengths), _ptr(rows), _ptr(jobs), cap, C.byref(nj))
            if rc != -4:   # APSE_ERR_CAPACITY: duplicated marker ids, retry with a larger job list
                break
            cap *= 4
        jobs = jobs[:nj.value]
    else:
        rc = lib.apse_sequence_scan(C.byref(cfg), F, M, _ptr(n), _ptr(ids), _ptr(corners), _ptr(rvec), _ptr(tvec),
                                    1 if rescale_tvec else 0, _ptr(lengths), None, None, 0, None)
    if rc != 0:
        raise ApseError(rc, "apse_sequence_scan failed")
    return lengths, rows, jobs


def _scan_chunk(lib, cfg, state, frame0, n, ids, corners, rvec, tvec, rescale_tvec, want_rows):
    n = np.ascontiguousarray(n, np.int32)
    F = int(n.shape[0])
    ids = np.ascontiguousarray(ids, np.int32).reshape(F, -1)
    M = int(ids.shape[1]) if F else 1
    corners = np.ascontiguousarray(corners, np.float32).reshape(F, M, 4, 2)
    rvec = np.ascontiguousarray(rvec, np.float64).reshape(F, M, 3)
    tvec = np.ascontiguousarray(tvec, np.float64).reshape(F, M, 3)
    lengths = np.empty(F, np.float64)
    args = (F, M, _ptr(n), _ptr(ids), _ptr(corners), _ptr(rvec), _ptr(tvec), 1 if rescale_tvec else 0, _ptr(lengths))
    if not want_rows:
        rc = lib.apse_sequence_scan_chunk(C.byref(cfg), _ptr(state), frame0, *args, None, None, 0, None)
        if rc != 0:
            raise ApseError(rc, "apse_sequence_scan_chunk failed")
        return lengths, None, None
    rows = np.zeros(F, SEQ_ROW_DTYPE)
    nj = C.c_int(0)
    cap = 4 * F + 64
    entry = state.copy()
    while True:
        jobs = np.empty(cap, SEQ_JOB_DTYPE)   # (the library clears every job it emits)
        rc = lib.apse_sequence_scan_chunk(C.byref(cfg), _ptr(state), frame0, *args, _ptr(rows), _ptr(jobs), cap, C.byref(nj))
        if rc != -4:
            break
        state[:] = entry   # duplicated marker ids: again from the chunk's entry state with a larger job list
        cap *= 4
    if rc != 0:
        raise ApseError(rc, "apse_sequence_scan_chunk failed")
    return lengths, rows, jobs[:nj.value]


def finish_chunk(state, rows, results):
    """apse_sequence_finish_chunk: the stale distance / LED values travel in the second pass's state."""
    rc = _lib.load().apse_sequence_finish_chunk(_ptr(state), len(rows), _ptr(rows) if len(rows) else None,
                                                _ptr(results) if len(results) else None, len(results))
    if rc != 0:
        raise ApseError(rc, "apse_sequence_finish_chunk failed")
    return rows


def run_jobs(engine, jobs, gray=None, frame0=0, results=None):
    """apse_sequence_jobs: all deferred projections of a sequence in one launch.  gray: [n,H,W] uint8 CUDA tensor holding the
    corrected gray frames frame0 .. frame0+n-1 (LED jobs of other frames keep valid = 0), or None."""
    lib = _lib.load()
    if results is None:
        results = np.zeros(len(jobs), SEQ_RESULT_DTYPE)
    if len(jobs) == 0:
        return results
    jobs = np.ascontiguousarray(jobs)
    (_, kp), (_, dp) = _lib.darr(engine.K), _lib.darr(engine.D)
    w, h = engine.size
    gptr, ng = (gray.data_ptr(), int(gray.shape[0])) if gray is not None else (None, 0)
    rc = lib.apse_sequence_jobs(engine.h, _ptr(jobs), len(jobs), gptr, int(frame0), ng, w, h, kp, dp, _ptr(results), engine._stream())
    if rc != 0:
        raise ApseError(rc, lib.apse_last_error(engine.h).decode())
    return results


def finish(rows, results):
    rc = _lib.load().apse_sequence_finish(len(rows), _ptr(rows), _ptr(results) if len(results) else None, len(results))
    if rc != 0:
        raise ApseError(rc, "apse_sequence_finish failed")
    return rows


def rows_to_csv(rows, header=True) -> str:
    """The text aruco_detect.py:131-139,146-185 writes for these rows."""
    rows = np.ascontiguousarray(rows)
    cap = 512 * (len(rows) + 2)
    buf = np.empty(cap, np.uint8)   # uninitialised: only the bytes the library reports are read
    nb = _lib.load().apse_sequence_csv(_ptr(rows), len(rows), 1 if header else 0, _ptr(buf), cap)
    if nb < 0:
        raise ApseError(int(nb), "apse_sequence_csv failed")
    return buf[:nb].tobytes().decode("ascii")


def rows_to_dicts(rows):
    """Rows in the shape postpass.SequencePostPass.step returns (for comparisons in the tests)."""
    out = []
    for r in rows:
        d = {"frame_ID": int(r["frame_id"]), "ID_4_detected": int(r["detected"][3])}
        if r["host_fields"]:
            d.update(markerLength=float(r["marker_length"]), leds_ID=int(r["leds"]), UAV_altitude=float(r["altitude"]),
                     fov_width=float(r["fov_width"]), fov_height=float(r["fov_height"]))
        else:
            d.update(markerLength=0, leds_ID=0, UAV_altitude=0, fov_width=0, fov_height=0)
        for v in (1, 2, 3):
            if r["detected"][v - 1]:
                d[f"ID_{v}_detected"] = 1
                d[f"distance_veh{v}_aruco"] = float(r["dist_aruco"][v - 1])
                d[f"distance_veh{v}_aruco_bbox"] = float(r["dist_bbox"][v - 1])
            else:
                d[f"ID_{v}_detected"] = d[f"distance_veh{v}_aruco"] = d[f"distance_veh{v}_aruco_bbox"] = 0
        out.append(d)
    return out


def _host_copy(engine, tensors):
    """One D2H transfer of several small device tensors: packed into one byte buffer on the device, copied into a cached
    pinned buffer, returned as numpy views (valid until the next call on this engine)."""
    torch = engine.torch
    flat = [t.contiguous().view(torch.uint8).reshape(-1) for t in tensors]
    sizes = [int(f.numel()) for f in flat]
    offs = np.concatenate([[0], np.cumsum([(s + 15) // 16 * 16 for s in sizes])]).astype(np.int64)
    total = int(offs[-1])
    pin = getattr(engine, "_seq_pin", None)
    if pin is None or pin.numel() < total:
        pin = engine._seq_pin = torch.empty(max(total, 1 << 20), dtype=torch.uint8).pin_memory()
    dev = torch.empty(max(total, 1), dtype=torch.uint8, device=engine.tdev)
    for f, o, s in zip(flat, offs, sizes):
        dev[int(o):int(o) + s] = f
    pin[:total].copy_(dev[:total], non_blocking=True)
    torch.cuda.current_stream(engine.tdev).synchronize()
    host = pin.numpy()
    return [host[int(o):int(o) + s].view(_NP[t.dtype]).reshape(tuple(t.shape)) for t, o, s in zip(tensors, offs, sizes)]


_NP = {}


def _np_types(torch):
    if not _NP:
        _NP.update({torch.int32: np.int32, torch.float32: np.float32, torch.float64: np.float64, torch.uint8: np.uint8})


def postpass_device(engine, det, start_frame=1, leds=False, leds_threshold=None, gray=None, frame0=0, exchange=None, details=False):
    """Post-pass of one sequence whose per-frame results `det` (dict of CUDA tensors n [F], ids [F,M], corners [F,M,4,2],
    rvec / tvec [F,M,3], poses computed with the nominal marker length) are on this rank's device.  Returns the rows.
    Only the first max(n) marker slots of every frame travel to the host (one packed pinned transfer per direction).
    exchange(jobs, results) (frame-sharded runs): lets the other ranks fill the LED jobs of the frames they own.
    details=True: dict(rows, jobs, results, n, ids, corners) instead of the rows alone."""
    torch = engine.torch
    _np_types(torch)
    F = int(det["n"].shape[0])
    w, h = engine.size
    cfg = seq_config(start_frame, 1, leds, leds_threshold, w, h)
    if F == 0:
        return np.zeros(0, SEQ_ROW_DTYPE)
    M = int(det["ids"].shape[1])
    m = max(1, min(M, int(det["n"].max().item())))           # marker slots in use (one 4-byte read-back)
    n, ids, corners, rvec, tvec = _host_copy(engine, [det["n"], det["ids"][:, :m], det["corners"][:, :m], det["rvec"][:, :m], det["tvec"][:, :m]])
    # pass 1: marker length of every frame from the nominal-length poses (tvec is linear in the marker length)
    lengths, _, _ = scan(cfg, n, ids, corners, rvec, tvec, rescale_tvec=True, want_rows=False)
    # pass 2: exact poses with those lengths, one launch for the sequence (aruco_detect.py:601 with that frame's markerLength)
    ml = torch.from_numpy(lengths.astype(np.float32)).to(engine.tdev)
    corners_d = det["corners"][:, :m].contiguous() if m < M else det["corners"]
    rv2, tv2 = engine.pose_frames(corners_d, det["n"], ml)
    n, ids, corners = n.copy(), ids.copy(), corners.copy()   # the pinned buffer is reused by the next transfer
    rv2h, tv2h = _host_copy(engine, [rv2, tv2])
    lengths2, rows, jobs = scan(cfg, n, ids, corners, rv2h, tv2h, rescale_tvec=False, want_rows=True)
    results = run_jobs(engine, jobs, gray=gray if leds else None, frame0=frame0)
    if exchange is not None:
        results = exchange(jobs, results)
    rows = finish(rows, results)
    if details:   # what the overlay renderer needs (render.py)
        return dict(rows=rows, jobs=jobs, results=results, n=n, ids=ids, corners=corners)
    return rows


class SequenceStream:
    """The post-pass of postpass_device for a sequence whose per-frame results arrive in chunks (frame order): every push runs
    both scans, the exact-pose launch and the projection jobs for its frames only, carrying the reference's globals in two
    apse_seq_state blocks -- the rows equal those of one postpass_device call over the whole sequence (tests/test_sequence_native.py,
    tests/test_gpu_pipeline.py).  The GPU work of a push goes to a high-priority stream of its own, so it does not queue behind
    pipeline batches that are still running.  LED read-out is not streamed (shard.run_sequence handles it)."""

    def __init__(self, engine, start_frame=1):
        self.engine = engine
        torch = engine.torch
        _np_types(torch)
        w, h = engine.size
        self.cfg = seq_config(start_frame, 1, False, None, w, h)
        self.s1, self.s2 = new_state(), new_state()
        self.frames = 0
        self.rows = []
        self.stream = torch.cuda.Stream(device=engine.tdev, priority=-1)

    def push(self, n, ids, corners, rvec, tvec, corners_dev, n_dev, ready=None):
        """Host arrays n [F], ids [F,M], corners [F,M,4,2], rvec / tvec [F,M,3] (nominal marker length) of the next F frames and the
        same corners / counts on the device ([F,M,4,2] float32, [F] int32, contiguous); ready: event the device copies wait for."""
        e = self.engine
        torch = e.torch
        F = int(len(n))
        if F == 0:
            return
        f0 = self.frames
        lengths, _, _ = scan(self.cfg, n, ids, corners, rvec, tvec, rescale_tvec=True, want_rows=False, state=self.s1, frame0=f0)
        with torch.cuda.stream(self.stream):
            if ready is not None:
                self.stream.wait_event(ready)
            ml = torch.from_numpy(lengths.astype(np.float32)).to(e.tdev)
            rv2, tv2 = e.pose_frames(corners_dev, n_dev, ml)
            rv2h, tv2h = _host_copy(e, [rv2, tv2])
            _, rows, jobs = scan(self.cfg, n, ids, corners, rv2h, tv2h, state=self.s2, frame0=f0)
            results = run_jobs(e, jobs)
        self.rows.append(finish_chunk(self.s2, rows, results))
        self.frames += F

    def result(self):
        return np.concatenate(self.rows) if self.rows else np.zeros(0, SEQ_ROW_DTYPE)
