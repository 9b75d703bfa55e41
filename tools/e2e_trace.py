"""Host-side timeline of the chunked grootgpu_align_batch pipeline (GROOTGPU_TRACE=1), for tuning on the GPU box.
   python tools/e2e_trace.py [n_reads]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from groot_b200 import api, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
L = 100
msa_dir = synth.unpack_db(os.path.join(ROOT, "data", "db", "arg-annot.90.tar"), "/tmp/groot_b200_db_%d" % os.getuid())
idx = api.Index.build(msa_dir=msa_dir, k=31, S=21, w=L)
blob, off = synth.synth_reads(n, L, synth.db_sequences(msa_dir), seed=42)
h_seq = torch.from_numpy(blob).pin_memory()
h_off = torch.from_numpy(off.view(np.int64)).pin_memory()
for chunk in (os.environ.get("GROOTGPU_CHUNK_READS", "1600000"), "2000000", "2500000"):
    os.environ["GROOTGPU_CHUNK_READS"] = chunk
    os.environ.pop("GROOTGPU_TRACE", None)
    for it in range(5):
        if it == 4 and chunk == "1600000":
            os.environ["GROOTGPU_TRACE"] = "1"
        t0 = time.perf_counter()
        raw = idx.map_reads_raw(h_seq.data_ptr(), h_off.data_ptr(), n, 0.99, project_on_device=True, compact=True)
        wall = (time.perf_counter() - t0) * 1e3
    print("chunk %s: %.2f ms wall (%.1f M reads/s), %.2f ms first copy-in to last copy-out, device %.2f ms" % (chunk, wall, n / wall / 1e3, raw.ms[0], raw.ms[1] + raw.ms[2] + raw.ms[3]))
