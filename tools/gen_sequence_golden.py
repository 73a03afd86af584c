"""Regenerates tests/golden/sequence_4k.json: the reference script's own CSV rows (tools/run_reference_script.py)
for a seeded synthetic 4K sequence.  Only seeds + CSV lines are stored; frames are re-rendered from the seeds."""
import json, os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cv2
from tools import synth, run_reference_script
from oracle import cv2_compat as C

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
d = C.Dictionary_get(C.DICT_4X4_50)
# (file, frames, seed, LED patterns of the host vehicle's strip per frame -- aruco_detect.py:338-373 reads them back)
for name, N, SEED, leds in (("sequence_4k.json", 8, 500, None),
                            ("sequence_4k_leds.json", 6, 520, [0b10110010, 0b01001101, 0b11111111, 0b00000000, 0b10000001, 0b01100110])):
    with tempfile.TemporaryDirectory() as tmp, tempfile.TemporaryDirectory() as tmp2:
        for k, f in enumerate(synth.make_sequence(d.bytesList, SEED, N, leds=leds)):
            cv2.imwrite(os.path.join(tmp, "image_%04d.png" % (k + 1)), f)
        csv = run_reference_script.run(tmp, os.path.join(tmp2, "out.csv"), os.path.join(root, "tests", "golden"))
    json.dump({"n_frames": N, "base_seed": SEED, "generator": "tools.synth.make_sequence", "leds": leds, "cv2": cv2.__version__,
               "csv": csv.splitlines()}, open(os.path.join(root, "tests", "golden", name), "w"), indent=1)
    print(name); print(csv)
