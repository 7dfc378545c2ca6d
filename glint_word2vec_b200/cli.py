"""Command line front-end: train from a text file, query and export models.

    python -m glint_word2vec_b200 train corpus.txt --out /data/model --vector-size 300 --num-servers 8
    python -m glint_word2vec_b200 synonyms /data/model wien --num 10
    python -m glint_word2vec_b200 export /data/model /data/local       # stock Word2VecModel layout (parquet)
    python -m glint_word2vec_b200 server --num-servers 8 --port 13370  # stand-alone shard-server group

The reference is driven from Spark code / spark-submit (README.md:26-67); this is the equivalent entry point for a
single box.  ``train`` uses the native text-file loader (``fitTextFile``), every hyper-parameter of the estimator is
available as ``--kebab-case`` option and engine options go through ``--config key=value`` (``parameterServerConfig``).
"""
from __future__ import annotations

import argparse
import json
import sys
from typing import List, Optional


def _kv(items: Optional[List[str]]) -> dict:
    out = {}
    for it in items or []:
        if "=" not in it:
            raise SystemExit(f"--config expects key=value, got {it!r}")
        k, v = it.split("=", 1)
        try:
            out[k] = json.loads(v)
        except json.JSONDecodeError:
            out[k] = v
    return out


def build_parser() -> argparse.ArgumentParser:
    ap = argparse.ArgumentParser(prog="glint_word2vec_b200", description=__doc__.split("\n\n")[0])
    sub = ap.add_subparsers(dest="cmd", required=True)

    tr = sub.add_parser("train", help="train on a text file (one sentence per line) and save the model")
    tr.add_argument("corpus")
    tr.add_argument("--out", required=True, help="model directory")
    tr.add_argument("--tokenizer", default="java", choices=["java", "whitespace"])
    tr.add_argument("--vector-size", type=int, default=100)
    tr.add_argument("--window-size", type=int, default=5)
    tr.add_argument("--step-size", type=float, default=0.01875)
    tr.add_argument("--max-iter", type=int, default=1)
    tr.add_argument("--min-count", type=int, default=5)
    tr.add_argument("--max-sentence-length", type=int, default=1000)
    tr.add_argument("--batch-size", type=int, default=50)
    tr.add_argument("--n", type=int, default=5, help="negatives per pair")
    tr.add_argument("--subsample-ratio", type=float, default=1e-6)
    tr.add_argument("--num-servers", type=int, default=1, help="column shards = GPUs (numParameterServers)")
    tr.add_argument("--server-host", default="", help="attach to a separate shard-server group ip[:port]")
    tr.add_argument("--num-partitions", type=int, default=1, help="asynchronous workers of the reference (staleness window)")
    tr.add_argument("--unigram-table-size", type=int, default=100_000_000, help="with --config sampler=table")
    tr.add_argument("--seed", type=int, default=None)
    tr.add_argument("--config", action="append", metavar="KEY=VALUE",
                    help="engine option (parameterServerConfig), e.g. subsample_mode=word2vec, device=cpu")

    sy = sub.add_parser("synonyms", help="nearest neighbours of words of a saved model")
    sy.add_argument("model")
    sy.add_argument("words", nargs="+")
    sy.add_argument("--num", type=int, default=10)
    sy.add_argument("--server-host", default="")
    sy.add_argument("--config", action="append", metavar="KEY=VALUE")

    ex = sub.add_parser("export", help="convert a saved model to the stock Word2VecModel layout (parquet)")
    ex.add_argument("model")
    ex.add_argument("out")
    ex.add_argument("--config", action="append", metavar="KEY=VALUE")

    sv = sub.add_parser("server", help="stand-alone shard-server group (cf. spark-submit --class glint.Main)")
    sv.add_argument("server_args", nargs=argparse.REMAINDER)
    return ap


def main(argv: Optional[List[str]] = None) -> int:
    args = build_parser().parse_args(argv)
    if args.cmd == "server":
        from .parallel import server
        server.main(args.server_args)
        return 0
    from . import ServerSideGlintWord2Vec, ServerSideGlintWord2VecModel
    if args.cmd == "train":
        est = ServerSideGlintWord2Vec(
            inputCol="sentence", outputCol="vector", vectorSize=args.vector_size, windowSize=args.window_size,
            stepSize=args.step_size, maxIter=args.max_iter, minCount=args.min_count,
            maxSentenceLength=args.max_sentence_length, batchSize=args.batch_size, n=args.n,
            subsampleRatio=args.subsample_ratio, numParameterServers=args.num_servers,
            numPartitions=args.num_partitions, unigramTableSize=args.unigram_table_size,
            parameterServerHost=args.server_host, parameterServerConfig=_kv(args.config))
        if args.seed is not None:
            est.setSeed(args.seed)
        model = est.fitTextFile(args.corpus, args.tokenizer)
        try:
            model.save(args.out)
            rep = model.trainingReport or {}
            print(json.dumps({"words": model.numWords, "vector_size": model.getVectorSize(), "model": args.out,
                              "pairs": rep.get("pairs"), "seconds": rep.get("seconds"),
                              "loss_per_pair": rep.get("loss_per_pair")}))
        finally:
            model.stop()
        return 0
    cfg = _kv(args.config) or None
    if args.cmd == "synonyms":
        model = ServerSideGlintWord2VecModel.load(args.model, args.server_host, cfg)
        try:
            for w in args.words:
                try:
                    res = model.findSynonymsArray(w, args.num)
                except KeyError:
                    print(json.dumps({"word": w, "error": "not in vocabulary"}, ensure_ascii=False))
                    continue
                print(json.dumps({"word": w, "synonyms": [[s, round(float(c), 6)] for s, c in res]},
                                 ensure_ascii=False))
        finally:
            model.stop()
        return 0
    if args.cmd == "export":
        model = ServerSideGlintWord2VecModel.load(args.model, "", cfg)
        try:
            model.toLocal().save(args.out)
            print(json.dumps({"exported": args.out, "words": model.numWords}))
        finally:
            model.stop()
        return 0
    return 2


if __name__ == "__main__":
    sys.exit(main())
