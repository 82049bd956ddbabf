import sys, os, numpy as np, torch, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from fluidnet_b200 import synth, tfluids
from gpu_backend import make_gpu_model
n = 128
flags = torch.from_numpy(synth.make_flags(n, n, n, True, nb=1, geometry=True)).cuda()
U = torch.from_numpy(synth.make_smooth_velocity(flags.cpu().numpy(), True, amp=2.0)).cuda()
p = torch.zeros_like(flags)
gm = make_gpu_model(synth.make_model(True))
lib = tfluids.context().lib
for variant in (0,):
  for _ in range(2):
    gm.forward((p, U, flags))
  torch.cuda.synchronize()
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for _ in range(10): gm.forward((p, U, flags))
  b.record(); torch.cuda.synchronize()
  print('  model:forward %.3f ms' % (a.elapsed_time(b) / 10))
buf = torch.zeros(8 * 1024, dtype=torch.int64, device="cuda")
lib.tfl_debug_conv_ts_counters(C.c_void_p(buf.data_ptr()))
gm.forward((p, U, flags))
torch.cuda.synchronize()
lib.tfl_debug_conv_ts_counters(C.c_void_p(0))
d = buf.cpu().numpy().reshape(-1, 8)
d = d[d[:, 6] != 0]
names = ["epi: edge exchange+barrier (sum)", "epi: all after loads (sum)", "issuer wait plane_full (sum)",
         "issuer wait d_empty (sum)", "issuer done at", "epilogue wg0 wait d_full (sum)", "cta end", "epi: wait d_full + tmem loads (sum)"]
print("ctas", len(d), "(counters accumulate over the 3 layer launches; 'at' values are of the last layer)")
for i, nm in enumerate(names):
    print("  %-42s median %10.0f" % (nm, np.median(d[:, i])))
