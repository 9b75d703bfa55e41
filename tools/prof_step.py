"""One bench step (C3 workload, reads resident in HBM, full result arrays) between cudaProfilerStart/Stop, for
   ncu --profile-from-start off ... python tools/prof_step.py [n_reads] [read_len] [compact]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from groot_b200 import api, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 100
compact = len(sys.argv) > 3 and sys.argv[3] == "compact"
db = "card.90" if L == 150 else "arg-annot.90"
msa_dir = synth.unpack_db(os.path.join(ROOT, "data", "db", db + ".tar"), "/tmp/groot_b200_db_%d" % os.getuid())
idx = api.Index.build(msa_dir=msa_dir, k=31, S=21, w=L)
blob, off = synth.synth_reads(n, L, synth.db_sequences(msa_dir), seed=42)
dev = torch.device("cuda", 0)
d_seq = torch.zeros(n * L + 64, dtype=torch.uint8, device=dev)
d_seq[: n * L].copy_(torch.from_numpy(blob))
d_off = torch.from_numpy(off.astype(np.uint32).view(np.int32)).to(dev)
for it in range(3):
    idx.map_reads_device(d_seq.data_ptr(), d_off.data_ptr(), n, L, L, 0.99, project_on_device=True, compact=compact)
idx.weights()
torch.cuda.synchronize()
torch.cuda.profiler.start()
raw = idx.map_reads_device(d_seq.data_ptr(), d_off.data_ptr(), n, L, L, 0.99, project_on_device=True, compact=compact)
idx.weights()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one step: %d reads, %d pairs, %d records, %d kernels" % (n, raw.n_pairs, raw.n_records, raw.kernel_launches))
