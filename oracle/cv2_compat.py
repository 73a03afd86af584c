"""TEST INFRASTRUCTURE ONLY -- compat shim that lets the reference's call sequence run on cv2 4.13.

The reference (aruco_detect.py) was written for opencv-contrib 4.2.0 (reference README.md:42); the image
ships opencv-python-headless 4.13.0, where the legacy free functions are gone.  This shim restores the
legacy names on top of the 4.13 objects so that the reference's OWN third-party implementation is the
checker for every parity test.  Nothing here may be imported by the product package
`apse_uav_b200`; only tests/, __graft_entry__.smoke() and bench.py's reference / cpu_baseline legs use it.

Each function cites the aruco_detect.py line whose call it serves.
"""
from __future__ import annotations

import json
import numpy as np
import cv2

aruco = cv2.aruco

DICT_4X4_50 = aruco.DICT_4X4_50
CORNER_REFINE_NONE = aruco.CORNER_REFINE_NONE
CORNER_REFINE_SUBPIX = aruco.CORNER_REFINE_SUBPIX
CORNER_REFINE_CONTOUR = aruco.CORNER_REFINE_CONTOUR
CORNER_REFINE_APRILTAG = aruco.CORNER_REFINE_APRILTAG


def Dictionary_get(dict_id):
    """aruco_detect.py:263  aruco.Dictionary_get(aruco.DICT_4X4_50)"""
    return aruco.getPredefinedDictionary(dict_id)


def DetectorParameters_create():
    """aruco_detect.py:191"""
    return aruco.DetectorParameters()


def detectMarkers(image, dictionary, corners=None, ids=None, parameters=None, rejectedImgPoints=None,
                  cameraMatrix=None, distCoeff=None):
    """aruco_detect.py:267.  cameraMatrix/distCoeff are unused by OpenCV unless CORNER_REFINE_CONTOUR."""
    if parameters is None:
        parameters = aruco.DetectorParameters()
    det = aruco.ArucoDetector(dictionary, parameters)
    return det.detectMarkers(image)


def estimatePoseSingleMarkers(corners, markerLength, cameraMatrix, distCoeffs):
    """aruco_detect.py:601.  Legacy contrib semantics: markerLength is a C++ float, object points are
    float32 (-L/2,L/2,0),(L/2,L/2,0),(L/2,-L/2,0),(-L/2,-L/2,0); one solvePnP(ITERATIVE) per marker."""
    L = np.float32(markerLength)
    h = np.float32(L / np.float32(2.0))
    obj = np.array([[-h, h, 0], [h, h, 0], [h, -h, 0], [-h, -h, 0]], np.float32).reshape(4, 1, 3)
    n = len(corners)
    rvecs = np.zeros((n, 1, 3), np.float64)
    tvecs = np.zeros((n, 1, 3), np.float64)
    for i in range(n):
        ok, r, t = cv2.solvePnP(obj, np.asarray(corners[i], np.float32).reshape(4, 1, 2), cameraMatrix,
                                distCoeffs)
        rvecs[i, 0] = r.ravel()
        tvecs[i, 0] = t.ravel()
    return rvecs, tvecs, obj


def drawMarker(dictionary, marker_id, side_pixels):
    return aruco.generateImageMarker(dictionary, marker_id, side_pixels)


def drawAxis(image, cameraMatrix, distCoeffs, rvec, tvec, length):
    """aruco_detect.py:617"""
    return cv2.drawFrameAxes(image, cameraMatrix, distCoeffs, rvec, tvec, length)


# ---------------------------------------------------------------------------------------------------------
# The reference's own configuration, restated from the cited lines (values only, no code copied).


def reference_parameters(refine=CORNER_REFINE_APRILTAG):
    """aruco_detect.py:190-203 (9 non-default values) + :266 (APRILTAG)."""
    p = aruco.DetectorParameters()
    p.minMarkerPerimeterRate = 0.01
    p.perspectiveRemovePixelPerCell = 8
    p.perspectiveRemoveIgnoredMarginPerCell = 0.33
    p.errorCorrectionRate = 2.0
    p.aprilTagMinClusterPixels = 100
    p.aprilTagMaxNmaxima = 5
    p.aprilTagCriticalRad = 20 * np.pi / 180
    p.aprilTagMaxLineFitMse = 1
    p.aprilTagMinWhiteBlackDiff = 100
    p.cornerRefinementMethod = refine
    return p


def read_camera_params(path):
    """aruco_detect.py:92-103"""
    with open(path, "r") as f:
        cam = json.load(f)
    return np.array(cam["mtx"]), np.array(cam["dist"])


def gamma_lut(gamma=2):
    """aruco_detect.py:537-540"""
    lut = np.empty((1, 256), np.uint8)
    for i in range(256):
        lut[0, i] = np.clip(pow(i / 255.0, gamma) * 255.0, 0, 255)
    return lut


def preprocess_frame(frame, mapx, mapy, lut):
    """aruco_detect.py:250-259"""
    frame = cv2.remap(frame, mapx, mapy, cv2.INTER_LINEAR)
    lab = cv2.cvtColor(frame, cv2.COLOR_RGB2LAB)
    lab[..., 0] = cv2.LUT(lab[..., 0], lut)
    return cv2.cvtColor(lab, cv2.COLOR_LAB2RGB)


def reference_chain(frame, mapx, mapy, lut, params, mtx, dist, marker_length=0.55, dict_id=DICT_4X4_50):
    """aruco_detect.py:589-601 for one frame: preprocess, gray, detect, pose."""
    corrected = preprocess_frame(frame, mapx, mapy, lut)
    gray = cv2.cvtColor(corrected, cv2.COLOR_BGR2GRAY)
    corners, ids, rejected = detectMarkers(gray, Dictionary_get(dict_id), parameters=params,
                                           cameraMatrix=mtx, distCoeff=dist)
    rvec = tvec = None
    if ids is not None and len(ids):
        rvec, tvec, _ = estimatePoseSingleMarkers(corners, marker_length, mtx, dist)
    return dict(corrected=corrected, gray=gray, corners=corners, ids=ids, rejected=rejected, rvec=rvec,
                tvec=tvec)
