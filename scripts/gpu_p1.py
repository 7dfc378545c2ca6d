"""Host-side profile of the public fit loop (why e2e_fit << e2e)."""
import cProfile, pstats, sys, time, io
import numpy as np, torch
sys.path.insert(0, ".")
from glint_word2vec_b200 import ServerSideGlintWord2Vec, _C
from glint_word2vec_b200.data.sampler import build_alias, zipf_counts
V, D, B, NS = 10_000_000, 512, 131072, 60
counts = zipf_counts(V, 200 * B, 1.0)
alias = build_alias(counts.astype(np.float64))
dev = torch.device("cuda", 0)
tok_dev = torch.empty(NS * B, dtype=torch.int32, device=dev)
_C.zipf_stream(torch.from_numpy(alias.packed()).to(dev), 2024, 0, tok_dev)
toks = tok_dev.cpu().numpy(); del tok_dev
offs = np.arange(0, NS * B + 1, 1000, dtype=np.int64)
est = ServerSideGlintWord2Vec(vectorSize=D, seed=1, stepSize=0.025, subsampleRatio=1e-4, numParameterServers=1, maxIter=3,
                              parameterServerConfig={"subsample_mode": "word2vec", "step_tokens": B})
pr = cProfile.Profile()
pr.enable()
m = est.fitEncoded(toks, offs, counts)
pr.disable()
rep = m.trainingReport
print("steps", rep["steps"], "seconds", rep["seconds"], "pairs/s", rep["pairs"] / rep["seconds"], "device_ms", rep["device_ms"])
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(35); print(s.getvalue()[:6000])
m.stop()
