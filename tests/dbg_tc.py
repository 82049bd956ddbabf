import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import oracle
from fluidnet_b200 import synth
from gpu_backend import make_gpu_model
from test_gpu_step import make_batch
orc = oracle.Oracle()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
batch = make_batch(n, True, plume=False)
mnp = synth.make_model(True)
p0 = (synth.make_density(batch["flags"], seed=77) - np.float32(0.5)) * np.float32(0.1)
wp, wU, wscale = oracle.model_forward(orc, mnp, p0, batch["UDiv"], batch["flags"])
gm = make_gpu_model(mnp)
for mode in ("fp32", "tf32", "tf32x3"):
    gm.set_mode(mode)
    gp, gU = gm.forward((torch.from_numpy(p0).cuda(), torch.from_numpy(batch["UDiv"]).cuda(),
                         torch.from_numpy(batch["flags"]).cuda()), return_scale=True)
    torch.cuda.synchronize()
    ep = np.abs(gp.cpu().numpy() - wp).max() / np.abs(wp).max()
    eU = np.abs(gU.cpu().numpy() - wU).max() / np.abs(wU).max()
    print("mode %-7s rel err p %.3e U %.3e (|p|max %.3g)" % (mode, ep, eU, np.abs(wp).max()), flush=True)
