"""Recipe that compiles the reference's own driver script into oracle/_ref/ (test infrastructure; outputs are git-ignored
build products that travel to the GPU box, the reference SOURCE is never copied into this repository).

/root/reference/aruco_detect.py is top-level procedural Python written for opencv-contrib 4.2 with hand-edited
constants.  It is compiled where it lies into two marshalled code objects:
  aruco_detect.cv2.bin   the unmodified algorithm text on the installed cv2 (+ the 4.13 legacy-name shim of
                         oracle/cv2_compat.py): generates the golden CSVs and serves as the CPU reference
  aruco_detect.swap.bin  the same text with ONLY the two import lines (aruco_detect.py:1-2) swapped to
                         `import apse_uav_b200 as cv2` / `from apse_uav_b200 import aruco`: the drop-in promise of north_star
In both, the user constants of aruco_detect.py:21,26,60-87 (GUI off, CSV on, image folder input, the three paths) are
redirected to globals the runner injects (tools/run_reference_script.py).
"""
import marshal
import os
import re
import sys

REF = "/root/reference/aruco_detect.py"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")


def patched_source(variant):
    src = open(REF).read()

    def sub(pat, rep, s):
        out, n = re.subn(pat, rep, s, count=1, flags=re.M)
        assert n == 1, pat
        return out
    src = sub(r"^showImage = True", "showImage = False", src)
    src = sub(r"^saveResults = False", "saveResults = True", src)
    src = sub(r"^useImages = False", "useImages = True", src)
    src = sub(r"^useVideo = True", "useVideo = False", src)
    src = sub(r'^path_camera_params = "your_path" \+ "cam_params.json"', "path_camera_params = __APSE_CAM_PARAMS__", src)
    src = sub(r'^    path_input_images = "your_path"', "    path_input_images = __APSE_IMAGE_DIR__", src)
    src = sub(r'^    path_output_results = "your_path"', "    path_output_results = __APSE_OUT_CSV__", src)
    if variant == "cv2":
        src = sub(r"^from cv2 import aruco", "aruco = __legacy_aruco__", src)
    elif variant == "swap":
        src = sub(r"^import cv2\s*$", "import apse_uav_b200 as cv2", src)
        src = sub(r"^from cv2 import aruco", "from apse_uav_b200 import aruco", src)
    else:
        raise ValueError(variant)
    return src


def build():
    if not os.path.exists(REF):
        return False
    os.makedirs(OUT, exist_ok=True)
    for variant in ("cv2", "swap"):
        code = compile(patched_source(variant), "aruco_detect.py[%s]" % variant, "exec")
        with open(os.path.join(OUT, "aruco_detect.%s.bin" % variant), "wb") as f:
            f.write(("%d.%d\n" % sys.version_info[:2]).encode())
            marshal.dump(code, f)
    return True


def load(variant):
    """Code object of the compiled script, or None when it was never built for this interpreter."""
    path = os.path.join(OUT, "aruco_detect.%s.bin" % variant)
    if not os.path.exists(path):
        if not build():
            return None
    with open(path, "rb") as f:
        ver = f.readline().decode().strip()
        if ver != "%d.%d" % sys.version_info[:2]:
            return None
        return marshal.load(f)


if __name__ == "__main__":
    print("built" if build() else "reference not present: nothing built")
