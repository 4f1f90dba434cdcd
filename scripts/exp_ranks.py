"""Multi-GPU experiment through the library's own NCCL communicator: python scripts/exp_ranks.py nranks waters
(forks one process per GPU; prints per-rank wall / device timings)."""
import json, os, sys, time, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import multiprocessing as mp

def worker(rank, world, path, key, q, steps):
    os.environ["VB_NCCL_KEY"] = key; os.environ["VB_SHARD_KEY"] = key
    os.environ.setdefault("VB_HOST_THREADS", str(max(2, (os.cpu_count() or 2) // world)))
    from valence_b200 import api
    e = api.Engine(path, device=rank)
    e.attach_comm(rank, world, key)
    out = []
    for i in range(steps):
        t = time.time(); r = e.energy(); dt = time.time() - t
        out.append({"wall_ms": 1e3 * dt, "tiles_ms": r["t_tiles_ms"], "1e_ms": r["t_1e_ms"], "dens_ms": r["t_density_ms"], "diag_ms": r["t_diag_ms"], "host_ms": r["t_host_setup_ms"], "E": r["energy"]})
    q.put((rank, out[-1]))
    e.close()

if __name__ == "__main__":
    world, n = int(sys.argv[1]), int(sys.argv[2])
    from valence_b200 import inputs
    p = tempfile.mktemp(suffix=".inp"); open(p, "w").write(inputs.write(inputs.water_cluster(n, tol=(10, 20, 10))))
    for env in ({}, {"VB_SHARD_SETUP": "0"}):
        os.environ.update(env)
        ctx = mp.get_context("spawn"); q = ctx.Queue()
        ps = [ctx.Process(target=worker, args=(r, world, p, f"exp{os.getpid()}_{len(env)}", q, 3)) for r in range(world)]
        [x.start() for x in ps]; res = sorted(q.get() for _ in ps); [x.join() for x in ps]
        print("==", world, "ranks", n, "waters", env)
        for r, o in res: print(r, json.dumps({k: (round(v, 1) if k != "E" else repr(v)) for k, v in o.items()}))
