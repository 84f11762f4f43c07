#!/usr/bin/env python3
"""Oracle loss trace of BASELINE configs[3] (1024x1024 pair, app.py weights) for the first nine closure evaluations.

One CPU evaluation at 1024^2 takes ~5 s, too slow for a unit test, so the trace is committed as a fixture
(tests/golden/config4_1024_losses.json).  Run from the repo root:  python tests/golden/make_config4_1024.py
The trace shows the reference's own behaviour on this input: unit-step L-BFGS (no line search) overshoots at the
seventh evaluation (0.348 -> 14.4); the CUDA path has to reproduce that, not "fix" it.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import nst_oracle as O  # noqa: E402

ws, bs = O.vgg19_random_weights(1234, 13)
content, style = O.synth_image(1024, 1024, 0), O.synth_image(1024, 1024, 1)
t = time.time()
r = O.run_oracle(ws, bs, content, [style], 10 ** 9, max_evals=9, **O.APP_WEIGHTS)
out = {"size": 1024, "content_seed": 0, "style_seed": 1, "total_loss": [float(v[0]) for v in r.losses],
       "how": "oracle/nst_oracle.py run_oracle(max_evals=9), torch CPU fp32", "seconds": round(time.time() - t, 1)}
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "config4_1024_losses.json"), "w") as f:
    json.dump(out, f, indent=1)
print(out)
