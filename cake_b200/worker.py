"""`python -m cake_b200.worker` — a cake worker endpoint on a B200, the Python twin of cake_b200/host/cake_worker
(reference: `cake run --mode worker --name N --topology T --address A`, worker.rs:79-597).

    python -m cake_b200.worker <model_dir> --topology topology.yml --name gpu1 [--address 0.0.0.0:10128]
                               [--cluster-key K] [--dtype bf16|f16] [--max-seq S] [--device 0]
    python -m cake_b200.worker <model_dir> --layers model.layers.16-31 ...
    python -m cake_b200.worker --topology topology.yml --name gpu1          # dry run: print the expanded layer list
    python -m cake_b200.worker --echo [--address 127.0.0.1:0]               # protocol only, no GPU

Only the safetensors shards that hold the worker's layers are mapped (utils/mod.rs:334-384).
"""
from __future__ import annotations

import argparse
import sys
from typing import List, Optional

from .parallel import expand_layers, load_topology
from .wire import RawTensor, WireWorker


class _Echo:
    """tests/protocol.rs MockWorker: echoes the activation."""
    dtype = "F16"

    def info(self):
        return "cpu", 0

    def clear_cache(self):
        pass

    def forward_ops(self, x: RawTensor, ops):
        for name, _, _ in ops:
            if name == "model.layers.99":
                raise LookupError(f"could not find layer {name}")
        return x


def worker_layers(topology_path: Optional[str], name: Optional[str], layers: Optional[str]) -> List[str]:
    """The layers this worker serves: --layers (topology syntax, comma separated) or the node `name` of the file."""
    if layers:
        return expand_layers([s for s in layers.split(",") if s])
    if not topology_path:
        raise SystemExit("error: need --layers or --topology/--name")
    topo = load_topology(topology_path)
    if name not in topo:
        raise SystemExit(f"error: could not find topology node for worker '{name}'")  # worker.rs:150-154
    return topo[name]["layers"]


def main(argv: Optional[List[str]] = None) -> int:
    ap = argparse.ArgumentParser(prog="python -m cake_b200.worker")
    ap.add_argument("model_dir", nargs="?")
    ap.add_argument("--topology")
    ap.add_argument("--name")
    ap.add_argument("--layers")
    ap.add_argument("--address", default="127.0.0.1:10128")
    ap.add_argument("--cluster-key")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "f16"])
    ap.add_argument("--max-seq", type=int, default=0)
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--echo", action="store_true")
    a = ap.parse_args(argv)
    host, port = a.address.rsplit(":", 1)
    if a.echo:
        backend = _Echo()
    else:
        names = worker_layers(a.topology, a.name, a.layers)
        if not a.model_dir:
            print("\n".join(names))
            return 0
        from .loader import open_model
        from .model import B200Transformer, Context
        from .wire import B200Backend
        cfg, vb = open_model(a.model_dir, {"self": {"layers": names}}, worker="self")
        ctx = Context(cfg, vb, a.dtype, device=a.device, max_seq=a.max_seq or None)
        backend = B200Backend(ctx, {n: B200Transformer.load(n, ctx) for n in names})
    w = WireWorker(backend, host, int(port), a.cluster_key)
    print(f"listening on {w.address}", flush=True)
    try:
        w.serve_forever()
    except KeyboardInterrupt:
        pass
    return 0


if __name__ == "__main__":
    sys.exit(main())
