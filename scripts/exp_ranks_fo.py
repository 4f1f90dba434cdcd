"""first_order_opt through the library's own NCCL communicator: python scripts/exp_ranks_fo.py nranks
Every rank's ham / ovl of (H2O)_16, orbitals 1 and 24, against the fast-oracle fixtures; then (H2O)_256 orbital 1 timed."""
import json, os, sys, time, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import multiprocessing as mp

def worker(rank, world, paths, key, q):
    os.environ["VB_NCCL_KEY"] = key; os.environ["VB_SHARD_KEY"] = key
    os.environ.setdefault("VB_HOST_THREADS", str(max(2, (os.cpu_count() or 2) // world)))
    import numpy as np
    from valence_b200 import api
    out = {}
    e = api.Engine(paths[16], device=rank); e.attach_comm(rank, world, key)
    e.energy()
    for iorb in (1, 24):
        fx = json.load(open(os.path.join(ROOT, "tests", "golden", "fast__w16_fo%d.json" % iorb)))
        H, S, _ = e.first_order(iorb)
        Hf, Sf = np.array(fx["ham"]), np.array(fx["ovl"])
        out["w16_fo%d" % iorb] = (float(abs(H - Hf).max() / max(1.0, abs(Hf).max())), float(abs(S - Sf).max() / max(1.0, abs(Sf).max())))
    e.close()
    e = api.Engine(paths[256], device=rank); e.attach_comm(rank, world, key + "b")
    e.energy()
    t = time.time(); r = e.energy(); H, S, st = e.first_order(1); out["w256_energy_plus_fo_s"] = time.time() - t
    out["w256_fo_kernel_ms"] = st["t_tiles_ms"]
    e.close()
    q.put((rank, out))

if __name__ == "__main__":
    world = int(sys.argv[1])
    from valence_b200 import inputs
    paths = {}
    for n in (16, 256):
        paths[n] = tempfile.mktemp(suffix=".inp"); open(paths[n], "w").write(inputs.write(inputs.water_cluster(n, tol=(10, 20, 10))))
    for env in ({},):
        os.environ.update(env)
        ctx = mp.get_context("spawn"); q = ctx.Queue()
        ps = [ctx.Process(target=worker, args=(r, world, paths, f"fo{os.getpid()}_{len(env)}", q)) for r in range(world)]
        [x.start() for x in ps]; res = sorted(q.get() for _ in ps); [x.join() for x in ps]
        print("==", world, "ranks", env, flush=True)
        for r, o in res: print(r, json.dumps(o), flush=True)
