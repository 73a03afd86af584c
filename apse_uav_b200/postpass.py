"""Host post-pass: the sequential marker logic of aruco_detect.py:598-782 over per-frame detection results.

Everything here is O(markers) per frame and order-dependent by definition (marker-length recurrence
:306-308,623,641; track gating on the previous frame :607-613,631,782), so it runs on the host over the few
KB per frame the GPU pipeline returns.  The heavy call inside it, cv2.projectPoints (:344,377,424,468), goes
through `project` (the CUDA kernel of pose.cu by default).

State that the reference keeps in module globals (markerLength, c?_prev, detected_ID_prev, msp?, leds, altitude,
dist_*) lives in SequencePostPass; values that the reference leaves stale between frames stay stale here too.
Configuration = the reference's defaults (useCentroidData False, sourceLidar False, N_avg 1, step_frame 1).
"""
from __future__ import annotations

import numpy as np

WIDTH, HEIGHT = 3840, 2160                      # aruco_detect.py:519
MARKER_LENGTH_ORG = 0.55                        # :520
MARKER_DIV = 1.2                                # :522
DIV = 1.013                                     # :523

# vehicle tables (:543-549, :583-586); index 0..2 = vehicles 1..3, index 3 = host (id 4)
VEH_COORDS = {1: np.float32([[0, 0.42, 0]]), 2: np.float32([[0, 0.59, 0]]), 3: np.float32([[0, 0.58, 0]]),
              4: np.float32([[0, 0.07, 0]])}
VEH4_LIDAR = np.float32([[-0.05, -0.80, 0]])
VEH_DIM = {1: [-1.95, 2.8, -0.9, 0.9], 2: [-1.68, 2.86, -0.87, 0.87], 3: [-1.32, 2.48, -0.86, 0.86],
           4: [-2.35, 2.49, -0.86, 0.86]}
LED_AXIS = np.float32([[-0.419, -0.42, 0], [-0.414, -0.305, 0], [-0.409, -0.19, 0], [-0.404, -0.07, 0],
                       [-0.399, 0.065, 0], [-0.393, 0.19, 0], [-0.388, 0.315, 0], [-0.382, 0.435, 0]])


def marker_length_correction(altitude):
    """:306-308"""
    return MARKER_LENGTH_ORG * (1 - 0.00057 * altitude / MARKER_DIV) / DIV


def yaw_zxy_deg(rvec):
    """First angle of scipy's Rotation.from_rotvec(rvec).as_euler('zxy', degrees=True) (:412-413): extrinsic
    z-x-y, R = Ry(c) Rx(b) Rz(a) -> a = atan2(R[1,0], R[1,1])."""
    r = np.asarray(rvec, np.float64).ravel()
    th = np.linalg.norm(r)
    if th < 1e-300:
        return 0.0
    k = r / th
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    R = np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * (Kx @ Kx)
    return float(np.degrees(np.arctan2(R[1, 0], R[1, 1])))


def marker_data(c, cx_prev, cy_prev, marker_length):
    """:271-288 -- c: (4,2) float32 corners.  float32 sums truncated by int() before the division, as written there."""
    cx = int(c[0][0] + c[1][0] + c[2][0] + c[3][0]) / 4
    cy = int(c[0][1] + c[1][1] + c[2][1] + c[3][1]) / 4
    side = lambda a, b: np.sqrt(np.power(c[a][0] - c[b][0], 2) + np.power(c[a][1] - c[b][1], 2))
    msp = (side(1, 0) + side(2, 1) + side(3, 2) + side(0, 3)) / 4
    if cx_prev is not None and cy_prev is not None:
        diff = np.sqrt(np.power(cx_prev - cx, 2) + np.power(cy_prev - cy, 2)) * marker_length / msp
    else:
        diff = 0
    return abs(cx), abs(cy), msp, diff


def scaled_bbox_dims(tvec, rvec, dim):
    """:406-420 (the drawing part is visualisation and omitted)."""
    alpha_h = np.arctan(tvec[0] / tvec[2])
    alpha_v = np.arctan(tvec[1] / tvec[2])
    yaw = round(yaw_zxy_deg(rvec), 2)
    alpha_h = alpha_h if yaw < 0 else -alpha_h
    alpha_v = alpha_v if yaw < 0 else -alpha_v
    return np.multiply(dim, [1 - alpha_h / 2, 1 + alpha_h / 2, 1 - alpha_v / 2, 1 + alpha_v / 2])


def bbox_points(dim):
    """:433-464 -- 20+20+8+8 points on the vehicle outline, z = 0, (x,y) = (width, length) axes."""
    o1 = np.linspace(dim[0], dim[1], 20)
    o2 = np.linspace(dim[2], dim[3], 8)
    pts = np.zeros((56, 3))
    pts[0:20, 0], pts[0:20, 1] = dim[2], o1
    pts[20:40, 0], pts[20:40, 1] = dim[3], o1
    pts[40:48, 0], pts[40:48, 1] = o2, dim[0]
    pts[48:56, 0], pts[48:56, 1] = o2, dim[1]
    return pts


def to_pixels(img_pts):
    """np.maximum(0, np.int32(imgpts).reshape(-1,2)) (:345,378,425,469): truncation toward zero, then clamp."""
    return np.maximum(0, np.int32(np.asarray(img_pts).reshape(-1, 2)))


class SequencePostPass:
    def __init__(self, project, start_frame=1, step_frame=1, led_mean=None, leds_threshold=None, led_sums=None):
        """project(obj (n,3), rvec (3,), tvec (3,)) -> (n,2) float64 image points (cv2.projectPoints with K, D).
        led_mean(x, y) -> mean of the 5x5 gray neighbourhood of the current frame (:356-358), or
        led_sums(xy (8,2) int) -> the 8 neighbourhood SUMS in one call (the GPU read-out); both None skips LEDs."""
        self.project, self.led_mean, self.leds_threshold, self.led_sums = project, led_mean, leds_threshold, led_sums
        self.start_frame = start_frame
        self.diff_max = 2 / 3 * step_frame * 2           # :524
        self.marker_length = MARKER_LENGTH_ORG           # :521
        self.detected_prev = [0, 0, 0, 0]                # :534
        self.prev_xy = {i: (0, 0) for i in (1, 2, 3, 4)}  # :535
        # stale-able per-vehicle values (module globals in the reference)
        self.cxy, self.msp, self.diff, self.size_corr, self.dim = {}, {}, {}, {}, {}
        self.leds, self.altitude = 0, 0.0
        self.dist = {1: (0.0, 0.0), 2: (0.0, 0.0), 3: (0.0, 0.0)}
        self.lidar_px = None

    # -- helpers ---------------------------------------------------------------------------------------------
    def _avg_size(self, msp):
        """:290-304 with N_avg = 1"""
        size_corr = np.sum(msp) / (msp * np.count_nonzero(msp))
        return size_corr, msp * size_corr

    def _leds(self, tvec, rvec, size_corr):
        """:338-373 (read-out only)"""
        if self.led_mean is None and self.led_sums is None:
            return self.leds
        px = to_pixels(self.project(LED_AXIS, rvec, tvec / size_corr))
        thr = max(190 + int(tvec[2] / MARKER_DIV), 240) if self.leds_threshold is None else self.leds_threshold
        if self.led_sums is not None:
            vals = np.asarray(self.led_sums(px[:8]), np.float64) / 25          # np.sum(np.sum(point)) / 25 (:358)
        else:
            vals = [self.led_mean(int(px[j][0]), int(px[j][1])) for j in range(8)]
        leds = 0
        for j in range(8):
            if vals[j] > thr:
                leds += 2 ** (7 - j)
        return leds

    # -- one frame -------------------------------------------------------------------------------------------
    def step(self, k, ids, corners, rvecs, tvecs):
        """ids: (n,) or (n,1) int array or None; corners (n,4,2) float32; rvecs/tvecs (n,3) float64 computed with
        self.marker_length as it was BEFORE this call (:601).  Returns the CSV fields of :146-185 as a dict."""
        detected = [0, 0, 0, 0]
        dims = {v: list(VEH_DIM[v]) for v in VEH_DIM}         # re-initialised every frame (:583-586)
        first = k == self.start_frame
        accepted = {}
        if ids is not None and len(ids):
            ids = np.array(ids, np.int64).reshape(-1).copy()
            corners = np.asarray(corners, np.float32).reshape(-1, 4, 2)
            rvecs = np.asarray(rvecs, np.float64).reshape(-1, 3)
            tvecs = np.asarray(tvecs, np.float64).reshape(-1, 3)
            for i in range(len(ids)):
                # the reference tests id 4, then the "[4] not in ids" altitude fallback, then ids 1, 2, 3 (:606-723)
                for vid in (4, 0, 1, 2, 3):
                    if vid == 0:
                        if 4 not in ids:                                            # :639-642
                            self.altitude = tvecs[i][2]
                            self.marker_length = marker_length_correction(self.altitude)
                            self.altitude = self.altitude / MARKER_DIV
                        continue
                    if ids[i] != vid:
                        continue
                    slot = vid - 1
                    px, py = (None, None) if first else self.prev_xy[vid]
                    cx, cy, msp, diff = marker_data(corners[i], px, py, self.marker_length)
                    self.cxy[vid], self.diff[vid] = (cx, cy), diff
                    if self.detected_prev[slot] == 0:                               # new marker or false positive
                        detected[slot] = 1
                        self.prev_xy[vid] = (cx, cy)
                    if (self.detected_prev[slot] == 1 and diff < self.diff_max) or first:
                        detected[slot] = 1
                        if vid == 4:
                            self.altitude = tvecs[i][2]                              # :622-624
                            self.marker_length = marker_length_correction(self.altitude)
                            self.altitude = self.altitude / MARKER_DIV
                        self.size_corr[vid], self.msp[vid] = self._avg_size(msp)
                        if vid == 4:
                            self.leds = self._leds(tvecs[i], rvecs[i], self.size_corr[4])
                            self.lidar_px = to_pixels(self.project(VEH4_LIDAR, rvecs[i], tvecs[i] / self.size_corr[4]))
                        self.prev_xy[vid] = (cx, cy)
                        dims[vid] = scaled_bbox_dims(tvecs[i], rvecs[i], dims[vid])
                        accepted[vid] = i
                    else:
                        ids[i] = -1                                                  # :637,669,696,723
            # distances from the host marker to every vehicle still carrying its id (:729-780)
            for i in range(len(ids)):
                if ids[i] != 4:
                    continue
                for j in range(len(ids)):
                    v = int(ids[j])
                    if v not in (1, 2, 3):
                        continue
                    if (self.detected_prev[v - 1] == 1 and self.diff[v] < self.diff_max) or first:
                        src = np.float32([[self.cxy[4][0], self.cxy[4][1]]])
                        px = to_pixels(self.project(bbox_points(dims[v]), rvecs[j], tvecs[j] / self.size_corr[v]))
                        best, idx = np.inf, 0
                        for t in range(len(px)):
                            dd = np.sqrt(pow(src[0][0] - px[t][0], 2) + pow(src[0][1] - px[t][1], 2))
                            if dd < best:
                                best, idx = dd, t
                        point = px[idx]
                        tgt = np.float32([[self.cxy[v][0], self.cxy[v][1]]])
                        d_aruco = np.sqrt((src[0][0] - tgt[0][0]) * (src[0][0] - tgt[0][0]) +
                                          (src[0][1] - tgt[0][1]) * (src[0][1] - tgt[0][1]))
                        d_bbox = np.sqrt((src[0][0] - point[0]) * (src[0][0] - point[0]) +
                                         (src[0][1] - point[1]) * (src[0][1] - point[1]))
                        scale = self.marker_length / ((self.msp[4] + self.msp[v]) / 2)
                        self.dist[v] = (float(d_aruco * scale), float(d_bbox * scale))
            self.detected_prev = detected                                            # :782 (inside the if)
        row = {"frame_ID": k, "ID_4_detected": detected[3]}
        if detected[3] == 1 and 4 in self.msp:
            row.update(markerLength=round(self.marker_length, 5), leds_ID=self.leds, UAV_altitude=round(float(self.altitude), 2),
                       fov_width=round(float(WIDTH * self.marker_length / self.msp[4]), 2),
                       fov_height=round(float(HEIGHT * self.marker_length / self.msp[4]), 2))
        else:
            row.update(markerLength=0, leds_ID=0, UAV_altitude=0, fov_width=0, fov_height=0)
        for v in (1, 2, 3):
            if detected[v - 1] == 1:
                row[f"ID_{v}_detected"] = 1
                row[f"distance_veh{v}_aruco"] = round(self.dist[v][0], 3)
                row[f"distance_veh{v}_aruco_bbox"] = round(self.dist[v][1], 3)
            else:
                row[f"ID_{v}_detected"] = row[f"distance_veh{v}_aruco"] = row[f"distance_veh{v}_aruco_bbox"] = 0
        return row


CSV_HEADER = ("frame_ID ,ID_4_detected ,markerLength ,leds_ID ,UAV_altitude ,fov_width ,fov_height ,"
              "ID_1_detected ,distance_veh1_aruco ,distance_veh1_aruco_bbox ,"
              "ID_2_detected ,distance_veh2_aruco ,distance_veh2_aruco_bbox ,"
              "ID_3_detected ,distance_veh3_aruco ,distance_veh3_aruco_bbox ,")  # :136-139
CSV_FIELDS = ["frame_ID", "ID_4_detected", "markerLength", "leds_ID", "UAV_altitude", "fov_width", "fov_height",
              "ID_1_detected", "distance_veh1_aruco", "distance_veh1_aruco_bbox",
              "ID_2_detected", "distance_veh2_aruco", "distance_veh2_aruco_bbox",
              "ID_3_detected", "distance_veh3_aruco", "distance_veh3_aruco_bbox"]


def csv_line(row):
    """:146-185"""
    return ",".join(str(row[f]) for f in CSV_FIELDS)
