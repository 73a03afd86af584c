"""Regenerates tests/golden/sequence_4k.json: the reference script's own CSV rows (tools/run_reference_script.py)
for a seeded synthetic 4K sequence.  Only seeds + CSV lines are stored; frames are re-rendered from the seeds."""
import json, os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cv2
from tools import synth, run_reference_script
from oracle import cv2_compat as C

N, SEED = 8, 500
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
d = C.Dictionary_get(C.DICT_4X4_50)
with tempfile.TemporaryDirectory() as tmp, tempfile.TemporaryDirectory() as tmp2:
    for k, f in enumerate(synth.make_sequence(d.bytesList, SEED, N)):
        cv2.imwrite(os.path.join(tmp, "image_%04d.png" % (k + 1)), f)
    csv = run_reference_script.run(tmp, os.path.join(tmp2, "out.csv"), os.path.join(root, "tests", "golden"))
json.dump({"n_frames": N, "base_seed": SEED, "generator": "tools.synth.make_sequence", "cv2": cv2.__version__, "csv": csv.splitlines()},
          open(os.path.join(root, "tests", "golden", "sequence_4k.json"), "w"), indent=1)
print(csv)
