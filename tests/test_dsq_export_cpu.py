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
