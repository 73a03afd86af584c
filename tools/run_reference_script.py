"""Runs the reference's own driver script, aruco_detect.py, on a folder of frames (test tooling).

The script is compiled from where it lies by oracle/build_ref.py (user constants of aruco_detect.py:21,26,60-87
redirected; nothing else touched) and executed here:
  module="cv2"   on the installed cv2 through the 4.13 legacy-name shim (oracle/cv2_compat.py) -> golden CSVs
  module="apse"  with only the two import lines swapped to apse_uav_b200 -> the import-swap promise, on a GPU
Returns the CSV text of aruco_detect.py:125-185.  Needs cv2 (frame I/O and drawing are passed through to it).
"""
import os, sys, types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def available(module="cv2"):
    from oracle import build_ref
    return build_ref.load("cv2" if module == "cv2" else "swap") is not None


def run(image_dir, out_csv, cam_params_dir, module="cv2"):
    from oracle import build_ref
    code = build_ref.load("cv2" if module == "cv2" else "swap")
    if code is None:
        raise RuntimeError("oracle/_ref/aruco_detect.*.bin missing: run `python oracle/build_ref.py` where /root/reference exists")
    g = {"__name__": "__reference_script__",
         "__APSE_CAM_PARAMS__": os.path.join(cam_params_dir, "cam_params.json"),
         "__APSE_IMAGE_DIR__": image_dir.rstrip("/") + "/",
         "__APSE_OUT_CSV__": out_csv}
    if module == "cv2":
        import cv2
        from oracle import cv2_compat
        shim = types.ModuleType("cv2_legacy_aruco")   # `from cv2 import aruco` resolves to the legacy-name shim
        for n in dir(cv2.aruco):
            if not n.startswith("__"):
                setattr(shim, n, getattr(cv2.aruco, n))
        for n in ("Dictionary_get", "DetectorParameters_create", "detectMarkers", "estimatePoseSingleMarkers", "drawAxis", "drawMarker"):
            setattr(shim, n, getattr(cv2_compat, n))
        g["__legacy_aruco__"] = shim
    exec(code, g)
    return open(out_csv).read()


if __name__ == "__main__":
    print(run(sys.argv[1], sys.argv[2], sys.argv[3], *(sys.argv[4:5])))
