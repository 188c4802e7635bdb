"""TEST INFRASTRUCTURE ONLY.  Writes tests/golden/metric_nms.json by running the UNMODIFIED reference metric
(/root/reference/utils/evaluate_utils.py: PostProcessRunner.nms_temporal / calculate_IoU) on seeded random segments.

    cd /root/reference && python /root/repo/oracle/make_metric_goldens.py      # needs /root/reference; CPU only
"""
import json
import os
import random
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "oracle", "shims"))
sys.path.insert(0, "/root/reference")
os.chdir("/root/reference")  # evaluate_utils.py:16 opens a relative path at import
from utils.evaluate_utils import PostProcessRunner  # noqa: E402

R = PostProcessRunner.__new__(PostProcessRunner)
random.seed(222)
cases = []
for _ in range(100):
    n = random.randint(1, 24)
    x1 = [round(random.random() * 0.8, 6) for _ in range(n)]
    x2 = [round(a + random.random() * 0.2 + 1e-3, 6) for a in x1]
    s = [round(random.random(), 6) for _ in range(n)]
    gt = sorted([round(random.random(), 6), round(random.random(), 6)])
    if gt[1] - gt[0] < 1e-3:
        gt[1] += 0.1
    cases.append({"x1": x1, "x2": x2, "s": s, "gt": gt, "picks": R.nms_temporal(x1, x2, s, 0.45),
                  "iou": [R.calculate_IoU((gt[0], gt[1]), (a, b)) for a, b in zip(x1, x2)]})
json.dump(cases, open(os.path.join(REPO, "tests", "golden", "metric_nms.json"), "w"))
print("wrote", len(cases), "cases")
