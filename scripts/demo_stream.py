"""Large-corpus demonstration of fitTextFile ("Next" #7): a synthetic text file is counted, encoded to
disk block by block and trained from a memory map; prints phase times, throughput and the peak host RSS.

    python scripts/demo_stream.py --tokens 200000000 --vocab 1000000 --dim 128 [--servers N]
"""
import argparse, json, os, resource, sys, tempfile, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def rss_mb():
    return resource.getrusage(resource.RUSAGE_SELF).ru_maxrss / 1024.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tokens", type=int, default=200_000_000)
    ap.add_argument("--vocab", type=int, default=1_000_000)
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--servers", type=int, default=1)
    ap.add_argument("--dir", default=None)
    args = ap.parse_args()
    from glint_word2vec_b200 import ServerSideGlintWord2Vec
    from glint_word2vec_b200.data.sampler import build_alias, zipf_counts, zipf_tokens
    work = tempfile.mkdtemp(prefix="gw2v-demo-", dir=args.dir)
    path = os.path.join(work, "corpus.txt")
    out = {"tokens": args.tokens, "vocab": args.vocab, "dim": args.dim, "servers": args.servers}
    t0 = time.time()
    alias = build_alias(zipf_counts(args.vocab, 10 ** 9, 1.0).astype(np.float64))
    chunk = 4_000_000
    with open(path, "w") as f:
        for i, lo in enumerate(range(0, args.tokens, chunk)):
            t = zipf_tokens(alias, min(chunk, args.tokens - lo), seed=1000 + i)
            lines = t[: len(t) // 1000 * 1000].reshape(-1, 1000)
            f.write("\n".join(" ".join(map(str, row)) for row in lines.tolist()))
            f.write("\n")
    out["generate_s"] = round(time.time() - t0, 1)
    out["text_bytes"] = os.path.getsize(path)
    out["rss_after_generate_mb"] = round(rss_mb())
    est = ServerSideGlintWord2Vec(vectorSize=args.dim, seed=1, minCount=5, numParameterServers=args.servers, subsampleRatio=1e-4,
                                  parameterServerConfig={"subsample_mode": "word2vec", "step_tokens": 131072,
                                                         "stream_threshold_bytes": 0, "scratch_dir": work})
    t1 = time.time()
    model = est.fitTextFile(path)
    out["fit_total_s"] = round(time.time() - t1, 1)
    rep = model.trainingReport
    out["train_s"] = round(rep["seconds"], 2)
    out["steps"] = rep["steps"]
    out["words_trained"] = rep["words"]
    out["pairs_per_s"] = round(rep["pairs"] / rep["seconds"])
    out["words_per_s"] = round(rep["words"] / rep["seconds"])
    out["loss_per_pair"] = round(rep["loss_per_pair"], 4)
    out["vocab_kept"] = model.numWords
    out["encoded_bytes"] = 4 * rep["words"]
    out["peak_rss_mb"] = round(rss_mb())
    # ru_maxrss counts file-backed pages of the memory-mapped text and token files too; the anonymous part is what the
    # process really holds
    for line in open("/proc/self/status"):
        if line.startswith(("RssAnon", "RssFile", "VmHWM")):
            out[line.split(":")[0].lower() + "_mb"] = round(int(line.split()[1]) / 1024)
    model.stop()
    print(json.dumps(out), flush=True)
    import shutil
    shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()
