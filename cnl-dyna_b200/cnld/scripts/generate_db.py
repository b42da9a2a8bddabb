"""
cnld.scripts.generate_db -- patch-to-patch / patch-to-node frequency-response database of an array
(mirror of /root/reference/cnld/scripts/generate_db.py).  Same Config defaults (:256-277), same
database contents, same resume behaviour (progress table, :186-225), same post-processing to
impulse responses (:143-164).  The per-frequency `process` jobs of the reference's process pool are
replaced by device-resident sweeps over blocks of frequencies (cnld.sweep), one process per GPU.
"""
import os

import numpy as np

from .. import abstract, database, impulse_response
from ..sweep import FrequencySweep, SweepSetup

_Config = dict(freqs=(0, 50e6, 200e3), sound_speed=1500., fluid_rho=1000., array_config='array.json',
               mesh_refn=11, use_fluid=True, format='FullFormat', aprx='paca', basis='linear', admis='2',
               eta=0.8, eps=1e-12, m=4, clf=16, eps_aca=1e-2, rk=0, q_reg=2, q_sing=4, strict=True,
               freq_interp=2)
Config = abstract.register_type('Config', _Config)


def postprocess(file, interp):
    """Frequency responses -> patch-to-patch impulse responses (generate_db.py:143-164)."""
    freqs, ppfr = database.read_patch_to_patch_freq_resp(file)
    t, ppir = impulse_response.fft_to_fir(freqs, ppfr, interp=interp, axis=-1, use_kkr=True)
    src, dst, times = np.meshgrid(np.arange(ppir.shape[0]), np.arange(ppir.shape[1]), t, indexing='ij')
    database.append_patch_to_patch_imp_resp(file, source_patch=src.ravel(), dest_patch=dst.ravel(),
                                            time=times.ravel(), displacement=ppir.ravel())


def main(cfg, file, write_over=False, array=None, device=0, block=12, store_nodes=True, nstreams=3):
    """Run (or resume) the sweep described by `cfg` and write `file`."""
    if not cfg.use_fluid:
        raise NotImplementedError("use_fluid=False is marked 'doesn't work right yet' in the reference "
                                  "(generate_db.py:93) and is not on the accelerated path")
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
    sw = FrequencySweep(setup, device=device, nstreams=nstreams)
    # Single writer thread: block b is written (one transaction, then its progress
    # rows) while block b+1 is on the GPU.  The queue holds one block, so at most two blocks of node
    # displacements are alive; a writer failure stops the sweep at the next block.
    import queue
    import threading
    q = queue.Queue(maxsize=1)
    failure = []

    def writer():
        while True:
            item = q.get()
            if item is None:
                return
            if failure:
                continue                 # drain after an error so the producer never blocks
            try:
                ids, xp, xn, t = item
                database.append_frequency_block(file, freqs[ids], wavenums[ids], xp, xn, setup.vertices, t)
                database.update_progress(file, ids + 1)      # job ids are 1-based like the reference's
            except BaseException as e:   # noqa: BLE001 -- re-raised in the main thread
                failure.append(e)

    th = threading.Thread(target=writer, name='cnld-db-writer', daemon=True)
    th.start()
    try:
        for b0 in range(0, len(todo), block):
            if failure:
                break
            ids = todo[b0:b0 + block]
            xp, xn, t = sw.run(freqs[ids], want_nodes=store_nodes)
            q.put((ids, xp, xn, t))
    finally:
        q.put(None)
        th.join()
        sw.close()
    if failure:
        raise failure[0]
    postprocess(file, cfg.freq_interp)
