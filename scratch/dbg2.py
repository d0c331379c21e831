import sys, numpy as np, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/deepseek-ocr.rs_b200')
from oracle import preprocess as P
from tests.helpers import tiny_model
from dsocr.engine import load_model
cfg, ck, d = tiny_model("bf16")
page = P.synthetic_page(700, 1400, seed=7)
vi = P.prepare_vision_input(page, 1024, 640, True)
g = P.image_to_tensor(vi["global"])
names = ["sam.pos_added"] + [f"sam.block.{i}" for i in range(cfg.sam_depth)] + ["sam.neck_conv1", "sam.neck_conv2", "sam.net3", "clip.embeddings", "clip.pre_layernorm"] + [f"clip.layer.{i}" for i in range(cfg.clip_layers)] + ["global_pre", "global_post"]
def run(seq):
    eng = load_model(d + "/config.json", d + "/model.safetensors", None, 0, "bf16")
    eng.set_option("record_taps", 1)
    for what in seq:
        if what == "s": eng.vision_encode(P.image_to_tensor(P.synthetic_page(640, 640, 1)), None, None)
    rows = eng.vision_encode(g, None, None)
    taps = {n: eng.tap(n).copy() for n in names}
    eng.close()
    return rows, taps
r1, t1 = run([])
r2, t2 = run(["s"])
r3, t3 = run([])
for n in names:
    a, b, c = t1[n], t2[n], t3[n]
    d12 = np.abs(a - b); d13 = np.abs(a - c)
    w = d12.size // (a.size // (4096 if n.startswith("sam.") and "net3" not in n else (256 if ("net3" in n or "global" in n) else 257)))
    idx = np.unravel_index(d12.argmax(), (a.size // w, w)) if d12.max() > 0 else None
    print(f"{n:22s} fresh-vs-after640 max {d12.max():.4g} at {idx}   fresh-vs-fresh max {d13.max():.4g}")
