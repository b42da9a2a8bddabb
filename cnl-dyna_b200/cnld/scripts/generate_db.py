"""
cnld.scripts.generate_db -- patch-to-patch / patch-to-node frequency-response database of an array
(mirror of /root/reference/cnld/scripts/generate_db.py).  Same entry point `main(cfg, args)` (:167-172:
args.file, args.write_over, args.threads), same Config fields and defaults (:256-277, format 'HFormat'
included), same database contents, same resume behaviour (progress table, :186-225), same
post-processing to impulse responses (:143-164), same command line (util.script_parser2, util.py:466-506).

What changed is who does the work.  The reference forks one process per frequency (`process`, :30-130);
here every worker owns one GPU and pushes blocks of frequencies through the device-resident sweep:

  cfg.format 'FullFormat'  dense Z + unpivoted LU           (cnld_b200_sweep_run,  ZFullMatrix recipe)
  cfg.format 'HFormat'     ACA H-matrix + batched GMRES      (cnld_b200_hsweep_run, ZHMatrix recipe with
                           cfg.aprx / admis / eta / clf / eps_aca / strict; H-LU replaced by GMRES)

`args.threads` is the number of worker processes = GPUs (the reference: size of the process pool); the
parent process is the only database writer (the reference: a lock around every append), writing block b
while the workers compute block b+1, through the C bulk writer (cnld_b200_db_append_block).
"""
import argparse
import os
import queue
import threading

import numpy as np

from .. import _lib, abstract, database, impulse_response
from ..sweep import FrequencySweep, SweepSetup, shard_cyclic

_Config = dict(freqs=(0, 50e6, 200e3), sound_speed=1500., fluid_rho=1000., array_config='array.json',
               mesh_refn=11, use_fluid=True, format='HFormat', aprx='paca', basis='linear', admis='2',
               eta=0.8, eps=1e-12, m=4, clf=16, eps_aca=1e-2, rk=0, q_reg=2, q_sing=4, strict=True,
               freq_interp=2)
Config = abstract.register_type('Config', _Config)

GMRES = dict(tol=1e-6, restart=50, maxiter=500, batch=96)  # the Krylov stand-in for the reference's H-LU
LAST_TIMING = {}      # wall seconds of the stages of the last main() call (setup, sweep + writer, finish, postprocess)


def postprocess(file, interp, freqs=None, ppfr=None):
    """Frequency responses -> patch-to-patch impulse responses (generate_db.py:143-164).  `freqs`, `ppfr`
    [P(src), P(dst), nf]: the responses of a sweep that ran in this process (the same doubles the database
    holds); without them they are read back from the file, as the reference does."""
    if ppfr is None:
        freqs, ppfr = database.read_patch_to_patch_freq_resp(file)
    t, ppir = impulse_response.fft_to_fir(freqs, ppfr, interp=interp, axis=-1, use_kkr=True)
    # the rows the reference builds with meshgrid(source, dest, time) and appends through pandas
    database.append_patch_to_patch_imp_resp_block(file, t, ppir)


def _solver_kind(cfg):
    fmt = str(cfg.format).lower()
    if fmt in ('hformat', 'h'):
        if str(cfg.aprx).lower() not in ('paca', 'aca', 'hca', 'inter_row'):
            raise ValueError(f'unknown approximation {cfg.aprx!r}')
        if str(cfg.basis).lower() != 'linear':
            raise NotImplementedError("only basis='linear' is on the accelerated path (generate_db.py:265)")
        return 'h'
    if fmt in ('fullformat', 'full'):
        return 'dense'
    raise ValueError(f"cfg.format must be 'HFormat' or 'FullFormat', not {cfg.format!r}")


class _Solver:
    """One GPU: blocks of frequencies -> (x_patch, x_node | None, seconds)."""

    def __init__(self, setup, cfg, device, nstreams, ring=False):
        self.kind = _solver_kind(cfg)
        self._ring, self._turn, self._use_ring = None, 0, ring      # ring: the consumer is done with a block before the third next one
        if self.kind == 'dense':
            self.sw = FrequencySweep(setup, device=device, nstreams=nstreams)
        else:
            self.setup = setup
            self.ctx = _lib.Context(setup.vertices, setup.triangles, q_reg=setup.q_reg, q_sing=setup.q_sing, device=device)
            self.ctx.set_mbk(setup.K, setup.M, setup.B)
            self.ctx.set_loads(setup.F, setup.AVG, setup.on_boundary)
            self.H = _lib.HMatrix(self.ctx, clf=cfg.clf, eta=cfg.eta, admis=cfg.admis, strict=cfg.strict)
            # aprx 'inter_row' has no accuracy knob in the reference (order m): map it like the C layer does
            self.eps_aca = cfg.eps_aca if str(cfg.aprx).lower() != 'inter_row' else 10.0 ** -max(2, int(cfg.m))

    def run(self, freqs, want_nodes):
        if self.kind == 'dense':
            out = None
            if want_nodes and self._use_ring:
                # three page-locked result buffers in rotation (one being filled, one queued, one with the
                # writer): the node displacements land in them straight from the device
                nf, P, N = len(freqs), self.sw.setup.npatch, self.sw.setup.nnodes
                if self._ring is None or self._ring.shape[1] < nf:
                    self._ring = np.empty((3, nf, P, N), np.complex128)
                out = self._ring[self._turn % 3][:nf]
                self._turn += 1
            return self.sw.run(freqs, want_nodes=want_nodes, out_nodes=out)
        xp, xn, _, t = self.H.sweep(freqs, c=self.setup.c, rho=self.setup.rho, eps_aca=self.eps_aca,
                                    want_nodes=want_nodes, **GMRES)
        return xp, xn, t

    def close(self):
        if self.kind == 'dense':
            self.sw.close()
        else:
            self.H.close()
            self.ctx.close()


def _worker(rank, dev, cfg_json, setup_packed, freqs, ids, block, store_nodes, nstreams, out):
    """Worker process of the multi-GPU driver: owns GPU `dev`, solves its frequencies block by block."""
    try:
        os.environ['CNLD_B200_DEVICE'] = str(dev)
        cfg = abstract.loads(cfg_json)
        setup = SweepSetup.unpack(setup_packed)
        solver = _Solver(setup, cfg, dev, nstreams)
        try:
            for b0 in range(0, len(ids), block):
                sel = ids[b0:b0 + block]
                out.put((sel,) + tuple(solver.run(freqs[sel], store_nodes)))
        finally:
            solver.close()
        out.put(('done', rank))
    except BaseException as e:   # noqa: BLE001 -- reported to the parent, which re-raises
        import traceback
        out.put(('error', rank, f'{type(e).__name__}: {e}\n{traceback.format_exc()}'))


def main(cfg, args, write_over=None, array=None, device=0, block=12, store_nodes=True, nstreams=3, devices=None):
    """Run (or resume) the sweep described by `cfg`.  `args` is the parsed command line of the reference
    (args.file, args.write_over, args.threads) or, for convenience, just the database path."""
    if isinstance(args, (str, os.PathLike)):
        file, threads = os.fspath(args), None
        write_over = bool(write_over)
    else:
        file, threads = args.file, getattr(args, 'threads', None)
        write_over = bool(getattr(args, 'write_over', False)) if write_over is None else bool(write_over)
    if not cfg.use_fluid:
        raise NotImplementedError("use_fluid=False is marked 'doesn't work right yet' in the reference "
                                  "(generate_db.py:93) and is not on the accelerated path")
    kind = _solver_kind(cfg)
    import time
    t_start = time.perf_counter()
    LAST_TIMING.clear()
    f_start, f_stop, f_step = cfg.freqs
    freqs = np.arange(f_start, f_stop + f_step, f_step)
    wavenums = 2 * np.pi * freqs / cfg.sound_speed
    array = array if array is not None else abstract.load(cfg.array_config)
    setup = SweepSetup.from_abstract(array, cfg.mesh_refn, c=cfg.sound_speed, rho=cfg.fluid_rho,
                                     q_reg=cfg.q_reg, q_sing=cfg.q_sing)
    done = None
    if os.path.isfile(file) and not write_over:
        done, _ = database.get_progress(file)
        if np.all(done):
            return
    else:
        if os.path.isfile(file):
            os.remove(file)
        os.makedirs(os.path.dirname(os.path.abspath(file)), exist_ok=True)
        database.create_db(file, **cfg._asdict())
        database.create_progress_table(file, len(freqs))
        database.append_node(file, node_id=range(setup.nnodes), x=setup.vertices[:, 0], y=setup.vertices[:, 1],
                             z=setup.vertices[:, 2])
    todo = np.arange(len(freqs)) if done is None else np.nonzero(~done)[0]
    # devices: count (GPUs 0..n-1) or explicit list of device ids (the reference's -t: pool size)
    if devices is None:
        devices = int(threads) if threads else 1
    dev_ids = list(range(devices)) if isinstance(devices, int) else [int(d) for d in devices]
    dev_ids = dev_ids[:max(1, len(todo))] or [device]
    ndev = len(dev_ids)
    if kind == 'h':
        block = min(block, 2)        # an H-path frequency takes seconds: small blocks keep the resume points close
    database.begin_bulk(file)        # the freq-resp indexes are rebuilt once, after the load (database.finish_bulk)
    LAST_TIMING['setup'] = time.perf_counter() - t_start
    t_sweep = time.perf_counter()
    writer_busy = [0.0]
    block_done = []       # single-process path: seconds since the start of the sweep stage at which each block was queued
    # a sweep that computes every frequency in this process keeps the patch responses for the post-processing
    xp_all = np.zeros((len(freqs), setup.npatch, setup.npatch), np.complex128) if len(todo) == len(freqs) else None

    # Single writer: block b is written -- one transaction, data and progress rows --
    # while block b+1 is being computed.  The queue holds one block per producer; a writer failure stops
    # the sweep at the next block.
    q = queue.Queue(maxsize=max(1, ndev))
    failure = []

    def writer():
        while True:
            item = q.get()
            if item is None:
                return
            if failure:
                continue                 # drain after an error so the producers never block
            try:
                ids, xp, xn, t = item
                t_w = time.perf_counter()
                database.append_frequency_block(file, freqs[ids], wavenums[ids], xp, xn, setup.vertices, t,
                                                job_ids=ids + 1)     # job ids are 1-based like the reference's
                if xp_all is not None:
                    xp_all[ids] = xp
                writer_busy[0] += time.perf_counter() - t_w
            except BaseException as e:   # noqa: BLE001 -- re-raised in the main thread
                failure.append(e)

    th = threading.Thread(target=writer, name='cnld-db-writer', daemon=True)
    th.start()
    try:
        if ndev == 1:
            solver = _Solver(setup, cfg, dev_ids[0] if devices != 1 else device, nstreams, ring=True)
            try:
                for b0 in range(0, len(todo), block):
                    if failure:
                        break
                    ids = todo[b0:b0 + block]
                    q.put((ids,) + tuple(solver.run(freqs[ids], store_nodes)))
                    block_done.append(time.perf_counter() - t_sweep)
            finally:
                solver.close()
        else:
            _run_workers(cfg, setup, freqs, todo, dev_ids, block, store_nodes, nstreams, q, failure,
                         cyclic=(kind == 'h'))
    finally:
        q.put(None)
        th.join()
    if failure:
        raise failure[0]
    LAST_TIMING['sweep_and_writer'] = time.perf_counter() - t_sweep
    LAST_TIMING['writer_busy'] = writer_busy[0]
    if len(block_done) > 1:
        LAST_TIMING['first_block'] = block_done[0]       # includes creating the device context and its workspaces
        LAST_TIMING['steady_frequencies_per_s'] = (len(todo) - min(block, len(todo))) / (block_done[-1] - block_done[0])
    t_fin = time.perf_counter()
    database.finish_bulk(file)
    LAST_TIMING['finish_indexes'] = time.perf_counter() - t_fin
    t_pp = time.perf_counter()
    if xp_all is not None:
        postprocess(file, cfg.freq_interp, freqs, np.ascontiguousarray(np.transpose(xp_all, (1, 2, 0))))
    else:
        postprocess(file, cfg.freq_interp)
    LAST_TIMING['postprocess'] = time.perf_counter() - t_pp
    LAST_TIMING['frequencies'] = int(len(todo))


def _run_workers(cfg, setup, freqs, todo, dev_ids, block, store_nodes, nstreams, q, failure, cyclic):
    """One process per GPU (multiprocessing 'spawn', like the reference's Pool but long-lived): contiguous
    shards of the frequency list for the dense path, cyclic for the H path (cost grows with k)."""
    import multiprocessing as mp
    from ..sweep import shard
    ndev = len(dev_ids)
    ctxmp = mp.get_context('spawn')
    out = ctxmp.Queue(maxsize=2 * ndev)
    packed, cfg_json = setup.pack(), abstract.dumps(cfg)
    procs = []
    for r in range(ndev):
        ids = todo[shard_cyclic(len(todo), r, ndev)] if cyclic else todo[slice(*shard(len(todo), r, ndev))]
        p = ctxmp.Process(target=_worker, args=(r, dev_ids[r], cfg_json, packed, freqs, ids, block, store_nodes, nstreams, out),
                          daemon=True)
        p.start()
        procs.append(p)
    alive = ndev
    try:
        while alive:
            try:
                item = out.get(timeout=5.0)
            except queue.Empty:
                if any(p.exitcode not in (None, 0) for p in procs):
                    raise RuntimeError('a generate_db worker process died')
                continue
            tag = item[0] if isinstance(item[0], str) else None
            if tag == 'done':
                alive -= 1
            elif tag == 'error':
                raise RuntimeError(f'generate_db worker {item[1]} failed: {item[2]}')
            else:
                if failure:
                    break
                q.put(item)
    finally:
        for p in procs:
            if p.is_alive() and (alive or failure):
                p.terminate()
            p.join(timeout=30)


def script_parser():
    """The command line of util.script_parser2 (util.py:466-506): -g dump the default config, -s show it,
    `file` [-c config.json] [-t workers/GPUs] [-w] run."""
    def run(args):
        if args.show_config:
            print(Config())
            return
        if args.generate_config:
            abstract.dump(Config(), args.generate_config)
            return
        if args.file:
            cfg = Config()
            if args.config:
                loaded = abstract.load(args.config)
                cfg = Config(**(loaded._asdict() if hasattr(loaded, '_asdict') else dict(loaded)))
            return main(cfg, args)

    parser = argparse.ArgumentParser(description=__doc__.split('\n\n')[0])
    parser.add_argument('-g', '--generate-config')
    parser.add_argument('-s', '--show-config', action='store_true')
    parser.add_argument('file', nargs='?')
    parser.add_argument('-c', '--config')
    parser.add_argument('-t', '--threads', type=int, help='worker processes = GPUs')
    parser.add_argument('-w', '--write-over', action='store_true')
    parser.set_defaults(func=run)
    return parser


if __name__ == '__main__':
    _args = script_parser().parse_args()
    _args.func(_args)
