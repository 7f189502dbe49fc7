"""tests/golden/config1_sea.npz: BASELINE config 1 run by the UNMODIFIED reference on the CPU.

    python tests/golden/make_golden_config1.py        # build container only (needs /root/reference), ~5 min

The reference's own ``tools.infer.evaluate -> eval_performance -> evalSEA`` (tools/infer.py:136-155,
56-133; tools/worse_only.py) on its own ``UperNetForSemanticSegmentation("ConvNeXt-T_CVST", 21, None)``
with seed-0 weights, 2 x 512^2 synthetic images, eps 4/255, n_iter 10, all three SEA losses
(tests/cfg1.py holds the flow; the only patches are the RNG source of the random start and
``"cuda" -> "cpu"`` for the hard-coded device strings).  Stored: clean / per-loss stats, the
per-image accuracies the attacks returned, the argmax maps and labels (int8), worst-case aACC and mIoU.
"""
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ref_shims  # noqa: E402

ref_shims.install()
sys.path.insert(0, "/root/reference")
import tools.infer as TI  # noqa: E402
from semseg.models import UperNetForSemanticSegmentation  # noqa: E402

import cfg1  # noqa: E402

if __name__ == "__main__":
    t0 = time.time()
    model, x, y = cfg1.build_inputs(UperNetForSemanticSegmentation)
    w = torch.tensor(TI.VOC_WTS)
    with cfg1.cuda_means_cpu():
        res = cfg1.sea_flow(TI, model, x, y, w)
    out = {"l_outs": res["l_outs"].astype(np.int8), "y": y.numpy().astype(np.int8),
           "x_checksum": np.array([float(x.double().sum()), float(x[0, 0, 0, :4].double().sum())]), "worst_Acc": res["worst_Acc"],
           "worst_Acc_indiv": res["worst_Acc_indiv"], "final_miou": res["final_miou"],
           "torch_version": np.array(torch.__version__), "cpu_seconds": time.time() - t0}
    for k in ["clean"] + cfg1.LOSSES:
        out["stats__" + k] = np.array([res[k]["mAcc"], res[k]["aAcc"], res[k]["mIoU"]], dtype=np.float64)
    for k in cfg1.LOSSES:
        out["acc__" + k] = res["acc"][k]
    np.savez_compressed(os.path.join(HERE, "config1_sea.npz"), **out)
    print({k: (v if np.asarray(v).size < 8 else np.asarray(v).shape) for k, v in out.items()})
