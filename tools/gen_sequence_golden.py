"""Regenerates tests/golden/sequence_4k*.json: the reference script's own CSV rows (tools/run_reference_script.py, i.e. the
compiled text of /root/reference/aruco_detect.py on cv2) for seeded synthetic 4K sequences.  Only seeds, events and CSV
lines are stored; the tests re-render the frames from the seeds.
  sequence_4k.json         8 frames, steady tracks
  sequence_4k_leds.json    6 frames with a rendered LED strip (leds_ID column, aruco_detect.py:338-373)
  sequence_4k_events.json  64 frames: a vehicle vanishes and returns, a vehicle and the host jump further than DIFF_MAX,
                           the host vanishes (altitude fallback :639-642), one frame has no marker at all (:599)
"""
import json, os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cv2
from tools import synth, run_reference_script
from oracle import cv2_compat as C

EVENTS = {10: {"hide": [2]}, 11: {"hide": [2]}, 12: {"hide": [2]},
          20: {"jump": {1: (400, 250)}},
          30: {"hide": [4]}, 31: {"hide": [4]},
          40: {"hide": [1, 2, 3, 4]},
          50: {"jump": {4: (-350, 300)}},
          57: {"hide": [3], "jump": {2: (-300, -280)}}}

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
d = C.Dictionary_get(C.DICT_4X4_50)
only = sys.argv[1:]
# (file, frames, seed, LED patterns of the host vehicle's strip per frame -- aruco_detect.py:338-373 reads them back, events)
for name, N, SEED, leds, events in (("sequence_4k.json", 8, 500, None, None),
                                    ("sequence_4k_leds.json", 6, 520, [0b10110010, 0b01001101, 0b11111111, 0b00000000, 0b10000001, 0b01100110], None),
                                    ("sequence_4k_events.json", 64, 540, [0b10110010, 0b01001101, 0b11111111, 0b00000000, 0b10000001], EVENTS)):
    if only and name not in only:
        continue
    with tempfile.TemporaryDirectory() as tmp, tempfile.TemporaryDirectory() as tmp2:
        for k, f in enumerate(synth.make_sequence(d.bytesList, SEED, N, leds=leds, events=events)):
            cv2.imwrite(os.path.join(tmp, "image_%04d.png" % (k + 1)), f)
        csv = run_reference_script.run(tmp, os.path.join(tmp2, "out.csv"), os.path.join(root, "tests", "golden"))
    json.dump({"n_frames": N, "base_seed": SEED, "generator": "tools.synth.make_sequence", "leds": leds,
               "events": None if events is None else {str(k): {"hide": v.get("hide", []), "jump": {str(i): list(x) for i, x in v.get("jump", {}).items()}}
                                                       for k, v in events.items()},
               "cv2": cv2.__version__, "csv": csv.splitlines()}, open(os.path.join(root, "tests", "golden", name), "w"), indent=1)
    print(name); print(csv)
