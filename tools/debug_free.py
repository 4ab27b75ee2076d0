import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from conftest import OUT_KEYS, load_golden, regenerate_case
from oracle import parq_oracle as O
from parq_b200.decoder import DecoderEngine
dev = torch.device("cuda:0")
c = regenerate_case(load_golden("small"))
eng = DecoderEngine(c["sd"], dev)
args = [c[k].to(dev) for k in ("tokens", "camera", "T_cp", "T_wp", "T_wl")]
free = {k: v.clone() for k, v in eng.forward(*args, c["H"], c["W"], debug=True).items()}
outs = [{k: free[k][i].cpu() for k in OUT_KEYS} for i in range(8)]
refs = O.refs_from_outputs(outs, c["sd"])
forced = eng.forward(*args, c["H"], c["W"], forced_refs=refs.to(dev), debug=True)
torch.cuda.synchronize()
for i in range(8):
    print(i, "coord_pos equal", bool(torch.equal(forced["coord_pos"][i], free["coord_pos"][i])),
          {k: float((forced[k][i] - free[k][i]).abs().max()) for k in ("features", "decoder_out", "pred_logits", "center_unnormalized")})
