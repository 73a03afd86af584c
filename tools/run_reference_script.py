"""Runs the UNMODIFIED algorithm text of the reference's aruco_detect.py on a folder of frames (test tooling).

The script is top-level procedural code written for opencv-contrib 4.2 with hand-edited constants, so it is
executed from a patched COPY of its source text held in memory (the reference tree is read-only and never
copied into this repository): only the user constants of aruco_detect.py:21,26,60-87 change, and `cv2.aruco`
resolves to the 4.13 compat shim (oracle/cv2_compat.py).  Returns the CSV text of aruco_detect.py:125-185.
Requires /root/reference and cv2; used by tests/test_postpass.py and to regenerate tests/golden/sequence_*.json.
"""
import os, re, sys, types

REF = "/root/reference/aruco_detect.py"


def run(image_dir, out_csv, cam_params_dir):
    import cv2
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import cv2_compat
    src = open(REF).read()
    sub = lambda pat, rep, s: re.sub(pat, rep, s, count=1, flags=re.M)
    src = sub(r"^showImage = True", "showImage = False", src)
    src = sub(r"^saveResults = False", "saveResults = True", src)
    src = sub(r"^useImages = False", "useImages = True", src)
    src = sub(r"^useVideo = True", "useVideo = False", src)
    src = sub(r'^path_camera_params = "your_path" \+ "cam_params.json"', f'path_camera_params = {os.path.join(cam_params_dir, "cam_params.json")!r}', src)
    src = sub(r'^    path_input_images = "your_path"', f"    path_input_images = {image_dir.rstrip('/') + '/'!r}", src)
    src = sub(r'^    path_output_results = "your_path"', f"    path_output_results = {out_csv!r}", src)
    # `from cv2 import aruco` must resolve to the legacy-name shim
    shim = types.ModuleType("cv2_legacy_aruco")
    for n in dir(cv2.aruco):
        if not n.startswith("__"):
            setattr(shim, n, getattr(cv2.aruco, n))
    for n in ("Dictionary_get", "DetectorParameters_create", "detectMarkers", "estimatePoseSingleMarkers", "drawAxis", "drawMarker"):
        setattr(shim, n, getattr(cv2_compat, n))
    src = src.replace("from cv2 import aruco", "aruco = __legacy_aruco__", 1)
    g = {"__name__": "__reference_script__", "__legacy_aruco__": shim}
    exec(compile(src, REF, "exec"), g)
    return open(out_csv).read()


if __name__ == "__main__":
    print(run(sys.argv[1], sys.argv[2], sys.argv[3]))
