"""The product-side exporter (dsocr/export.py + the library writer) on the tiny checkpoint: same tensor list and dtype
assignment as the oracle's restatement of the adapter / dsq-cli chain, Q8_0 payloads byte-identical, K-quant payloads valid
and close, file accepted by the library's reader.  No GPU."""
import ctypes as C

import numpy as np
import pytest

from oracle import dsq
from tests.helpers import tiny_model
from tests.test_dsq_reader_cpu import Hdr, Rec, lib  # noqa: F401


@pytest.mark.parametrize("primary,name", [(dsq.Q8_0, "q8_0"), (dsq.Q4K, "q4k"), (dsq.Q6K, "q6k")])
def test_export_matches_oracle_assignment_and_bytes(lib, tmp_path, primary, name):
    from dsocr.export import export_snapshot

    cfg, ck, d = tiny_model("bf16")
    ref_path = str(tmp_path / "ref.dsq")
    assigned = dsq.write_model_snapshot(ref_path, cfg, ck, primary)
    written = export_snapshot(d + "/config.json", d + "/model.safetensors", str(tmp_path / "mine"), name)
    assert written == assigned                       # same tensors (order-insensitive) and dtype per tensor
    assert list(written) == [n for n, _, _ in dsq.linear_specs(cfg)]   # adapter order
    hdr = Hdr()
    assert lib.dsocr_dsq_inspect(str(tmp_path / "mine.dsq").encode(), C.byref(hdr), None, 0) == 0, lib.dsocr_last_error().decode()
    assert hdr.tensor_count == len(assigned) and hdr.default_qdtype == primary
    _, ref_recs, ref_data = dsq.read_snapshot(ref_path)
    _, my_recs, my_data = dsq.read_snapshot(str(tmp_path / "mine.dsq"))
    worst = 0.0
    for n, r in my_recs.items():
        rr = ref_recs[n]
        assert (r.out_dim, r.in_dim, r.q_dtype, r.q_len) == (rr.out_dim, rr.in_dim, rr.q_dtype, rr.q_len)
        mine = my_data[r.q_offset:r.q_offset + r.q_len]
        if r.q_dtype == dsq.Q8_0:
            assert mine == ref_data[rr.q_offset:rr.q_offset + rr.q_len], n
        elif n.endswith("layers.1.self_attn.q_proj.weight") or n.endswith("experts.0.gate_proj.weight"):
            w = ck[n].float().numpy()
            deq = dsq.dequantize(mine, r.q_dtype, r.out_dim, r.in_dim)
            worst = max(worst, float(np.sqrt(((deq - w) ** 2).mean()) / np.sqrt((w ** 2).mean())))
    assert worst < (0.09 if primary == dsq.Q4K else 0.03)


def test_adapter_discovers_projector_and_lm_head():
    """crates/dsq-models/tests/adapters.rs:4-36 (deepseek_adapter_discovers_projector_and_lm_head): the same config JSON."""
    from dsocr.export import linear_specs, select_dtype, Q4K, Q6K, Q8_0

    cfg = {"model_type": "deepseek_vl_v2", "hidden_size": 8, "intermediate_size": 16, "num_hidden_layers": 1, "num_attention_heads": 2,
           "num_key_value_heads": 2, "n_routed_experts": 0, "n_shared_experts": 0, "moe_layer_freq": 1, "first_k_dense_replace": 0,
           "lm_head": True, "vocab_size": 32, "projector_config": {"n_embed": 8, "input_dim": 4}}
    specs = linear_specs(cfg, include_projector=True)
    names = [s[0] for s in specs]
    assert "lm_head.weight" in names and "model.projector.layers.weight" in names
    assert specs[-1] == ("model.projector.layers.weight", 8, 4, "model.projector.layers.bias")
    assert names[:7] == [f"model.layers.0.self_attn.{n}_proj.weight" for n in "qkvo"] + \
        [f"model.layers.0.mlp.{n}_proj.weight" for n in ("gate", "up", "down")]   # n_routed_experts = 0 -> dense MLP
    assert "model.projector.layers.weight" not in [s[0] for s in linear_specs(cfg)]
    # recommend_dtype (deepseek_ocr.rs:141-154) + select_dtype (dsq-cli main.rs:953-1004)
    assert select_dtype("lm_head.weight", 1280, Q4K) == Q8_0 and select_dtype("lm_head.weight", 1280, Q8_0) == Q8_0
    assert select_dtype("model.layers.1.self_attn.q_proj.weight", 1280, Q6K) == Q6K
    assert select_dtype("model.layers.1.mlp.experts.0.down_proj.weight", 896, Q4K) == Q8_0   # 896 % 256 != 0 -> Q8_0
    assert select_dtype("x", 30, Q4K) is None                                               # -> float payload
