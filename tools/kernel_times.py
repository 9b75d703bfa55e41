"""Per-kernel device times of one batch (CUDA events inside libgrootgpu), for tuning on the GPU box.
   GROOTGPU_LIB=<variant .so> python tools/kernel_times.py [n_reads] [read_len]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from groot_b200 import api, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 100
db = "card.90" if L == 150 else "arg-annot.90"
msa_dir = synth.unpack_db(os.path.join(ROOT, "data", "db", db + ".tar"), "/tmp/groot_b200_db_%d" % os.getuid())
idx = api.Index.build(msa_dir=msa_dir, k=31, S=21, w=L)
blob, off = synth.synth_reads(n, L, synth.db_sequences(msa_dir), seed=42)
dev = torch.device("cuda", 0)
d_seq = torch.zeros(n * L + 64, dtype=torch.uint8, device=dev)
d_seq[: n * L].copy_(torch.from_numpy(blob))
d_off = torch.from_numpy(off.astype(np.uint32).view(np.int32)).to(dev)
for it in range(5):
    t0 = time.perf_counter()
    raw = idx.map_reads_device(d_seq.data_ptr(), d_off.data_ptr(), n, L, L, 0.99, project_on_device=True, compact=os.environ.get("KT_COMPACT") == "1")
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3
print("compact" if os.environ.get("KT_COMPACT") == "1" else "full", dict(zip(api.KERNEL_FAMILIES, ["%.3f" % v for v in list(raw.kernel_ms)[:8]])))
print("lib=%s n=%d L=%d seed=%.3f ms align=%.3f ms other=%.3f ms total_dev=%.3f ms wall=%.3f ms slow_path_pairs=%d pairs=%d records=%d"
      % (os.path.basename(api.LIB_PATH), n, L, raw.ms[1], raw.ms[2], raw.ms[3], raw.ms[0], wall, raw.slow_path_pairs, raw.n_pairs, raw.n_records))
