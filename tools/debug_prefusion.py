"""DenseNet-121 bf16: pre-activation fusion (BN + ReLU on the A tiles) against the unfused path -- must be bit-identical."""
import os
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
if len(sys.argv) > 1:  # child: run and dump
    from sykepic_b200 import engine, synth

    mdir = synth.write_model_dir(Path(sys.argv[1]) / "m", arch="densenet121", t=224, head=(256, 128), seed=3, border="mode",
                                 imagenet_normalization=False, randomize_bn=True, logit_gain=8.0)
    b = synth.synth_bin(1011, 40)
    eng = engine.Engine(mdir, precision="bf16", max_batch=32)
    rid, probs = eng.run_bin(b["adc_text"], b["roi_bytes"])
    np.save(sys.argv[2], probs)
    print("launches", eng.launches)
    sys.exit(0)
tmp = tempfile.mkdtemp()
outs = []
for tag, env in (("fused", {}), ("unfused", {"SPK_NO_PRE_FUSION": "1"})):
    o = os.path.join(tmp, tag + ".npy")
    r = subprocess.run([sys.executable, __file__, tmp, o], env=dict(os.environ, **env), capture_output=True, text=True)
    print(tag, r.stdout.strip()[-60:], r.stderr.strip()[-300:])
    outs.append(np.load(o))
print("max |fused - unfused| =", float(np.abs(outs[0] - outs[1]).max()), "identical:", bool(np.array_equal(outs[0], outs[1])))
