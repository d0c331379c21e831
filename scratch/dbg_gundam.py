import sys, numpy as np, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/deepseek-ocr.rs_b200')
from oracle import preprocess as P, vision as V
from tests.helpers import tiny_model, report
from dsocr.engine import load_model
cfg, ck, d = tiny_model("bf16")
oracle = V.VisionOracle(cfg, ck)
page = P.synthetic_page(700, 1400, seed=7)
vi = P.prepare_vision_input(page, 1024, 640, True)
g = P.image_to_tensor(vi["global"])
tiles = np.stack([P.image_to_tensor(t) for t in vi["tiles"]])
ref_g = oracle.encode(torch.from_numpy(g), None, None)
def run(tag, seq):
    eng = load_model(d + "/config.json", d + "/model.safetensors", None, 0, "bf16")
    for what in seq:
        if what == "g":
            rows = torch.from_numpy(eng.vision_encode(g, None, None)); report(tag + " global-only", rows, ref_g)
            bad = ((rows - ref_g).abs().max(1).values > 0.2).nonzero().flatten().tolist(); print("  bad rows", bad[:40], len(bad))
        elif what == "G":
            rows = torch.from_numpy(eng.vision_encode(g, tiles, vi["crop_shape"])); n_local = rows.shape[0] - 273
            report(tag + " gundam-global", rows[n_local:], ref_g)
            bad = ((rows[n_local:] - ref_g).abs().max(1).values > 0.2).nonzero().flatten().tolist(); print("  bad rows", bad[:40], len(bad))
        elif what == "s":
            eng.vision_encode(P.image_to_tensor(P.synthetic_page(640, 640, 1)), None, None)
    eng.close()
run("fresh", ["g"])
run("after640", ["s", "g"])
run("gundam-first", ["G"])
run("gundam-after-g", ["g", "G"])
