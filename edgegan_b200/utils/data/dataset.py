"""Dataset of the reference (`edgegan/utils/data/dataset.py:18-89`) plus a prefetching device feeder.

Directory layout (dataset.py:26-43): train, multi-class: <dataroot>/<name>/train/<class id>/*.{png,jpg} for class ids
0..num_classes-1; train, single class: <dataroot>/<name>/train/*.png; test: every png / jpg below
<dataroot>/<name>/test, sorted.  A sample is one [H, 2W, 3] picture: sketch on the left half, photo on the right.
`__getitem__(idx)` -> (images float32 [B, H, W, 3] in [-1, 1], z float64 [B, z_dim (+1: class id)], filenames) for
train, (images, filenames) for test; z comes from numpy's global generator like in the reference (dataset.py:72-73).
"""
from __future__ import annotations

import os
import queue
import threading
from concurrent.futures import ThreadPoolExecutor
from glob import glob
from pathlib import Path

import numpy as np

from ..utils import _TO_UNIT, get_image_bytes, get_image_fast


def _read_one(args, filename):
    ih, iw, oh, ow, crop, gray = args
    return get_image_fast(filename, input_height=ih, input_width=iw, resize_height=oh, resize_width=ow, crop=crop,
                          grayscale=gray)


def _read_many_bytes(args, filenames):
    """Worker-side decode: uint8 [n, H, W, C] (stretched + resized, before x / 127.5 - 1), or float32 when a file does
    not take the byte path (grayscale)."""
    ih, iw, oh, ow, crop, gray = args
    out = [get_image_bytes(f, ih, iw, oh, ow, crop, gray) for f in filenames]
    if all(o.dtype == np.uint8 for o in out):
        return np.stack(out)
    return np.stack([_TO_UNIT[o] if o.dtype == np.uint8 else o.astype(np.float32) for o in out])


def _read_many(args, filenames):
    """float32 [len(filenames), H, W, C]; top-level so that a process pool can run it."""
    return np.stack([_read_one(args, f) for f in filenames]).astype(np.float32, copy=False)


def extension_match_recursive(root, exts):
    """dataset.py:10-15."""
    result = []
    for ext in exts:
        result.extend(str(p) for p in Path(root).rglob(ext))
    return result


class Dataset():
    """Same constructor arguments, attributes (`data`, `size`, `batchsize`, ...) and error messages as the reference."""

    def __init__(self, dataroot, name, size, batchsize, config, num_classes=None, phase='train'):
        if phase not in ('train', 'test'):
            raise AssertionError(phase)
        self.batchsize, self.num_classes, self.config, self.phase = batchsize, num_classes, config, phase
        base = os.path.join(dataroot, name, phase)
        if phase == 'test':
            # every picture below the test directory, in sorted order
            where = base
            files = sorted(extension_match_recursive(base, ['*.png', '*.jpg']))
        elif num_classes is None:
            # single class: png files directly in the train directory
            where = os.path.join(base, '*.png')
            files = glob(where)
        else:
            # one directory per class id 0 .. num_classes-1 (png first, then jpg, class by class)
            files = []
            for cls in range(num_classes):
                for pattern in ('*.png', '*.jpg'):
                    where = os.path.join(base, str(cls), pattern)
                    files += glob(where)
        self.data = files
        if not files:
            raise Exception("[!] No data found in '" + where + "'")
        if len(files) < batchsize:
            raise Exception("[!] Entire dataset size is less than the configured batch_size")
        self._cap = size
        self.size = min(len(files), size)

    def shuffle(self, seed=None, rank=0, world=1):
        """dataset.py:53-54: shuffle in place with numpy's global generator.  Data-parallel training passes a `seed`
        shared by all ranks plus (rank, world): every rank then applies the SAME permutation to the full file list and
        keeps every world-th file, so the ranks see disjoint shards of one global epoch.  The permuted list is first
        truncated to a multiple of `world`: all shards have the SAME length, hence every rank runs the same number of
        batches and issues the same sequence of collectives (a shard one file longer could otherwise mean one more
        batch -- and one more gradient all-reduce -- on some ranks only)."""
        if seed is None:
            np.random.shuffle(self.data)
            return
        if not hasattr(self, "_all"):
            self._all = sorted(self.data)
        files = list(self._all)
        np.random.RandomState(seed).shuffle(files)
        files = files[:(len(files) // world) * world]
        self.data = files[rank::world]
        cap = self._cap if self._cap == float("inf") else int(self._cap) // world    # `train_size` caps the GLOBAL epoch
        self.size = int(min(len(files) // world, cap))

    def __len__(self):
        return self.size // self.batchsize

    def _read(self, filename):
        return _read_one(self._read_args(), filename)

    def _read_args(self):
        c = self.config
        return (c['input_height'], c['input_width'], c['output_height'], c['output_width'], c['crop'], c['grayscale'])

    @staticmethod
    def class_of(path):
        """dataset.py:76-79: the name of the directory holding the file."""
        end = path.rfind("/")
        start = path.rfind("/", 0, end)
        return int(path[start + 1:end])

    def load_batch(self, idx, pool=None, chunks=1):
        """`__getitem__` with the per-file decode optionally spread over a process / thread pool (`chunks` tasks)."""
        filenames = self.data[idx * self.batchsize:(idx + 1) * self.batchsize]
        if pool is None:
            batch_images = _read_many(self._read_args(), filenames)
        else:
            step = -(-len(filenames) // max(1, chunks))
            parts = [filenames[i:i + step] for i in range(0, len(filenames), step)]
            args = self._read_args()
            batch_images = np.concatenate(list(pool.map(_read_many, [args] * len(parts), parts)), axis=0)
        if self.phase == 'test':
            assert batch_images.shape[0] == len(filenames)
            return batch_images, filenames
        batch_z = np.random.normal(size=(self.batchsize, self.config['z_dim']))
        if self.num_classes is not None:
            classes = np.array([self.class_of(f) for f in filenames]).reshape((self.batchsize, 1))
            batch_z = np.concatenate((batch_z, classes), axis=1)
        return batch_images, batch_z, filenames

    def submit_batch(self, idx, pool, chunks=1):
        """Start decoding batch `idx` on `pool` -> futures (see finish_batch)."""
        filenames = self.data[idx * self.batchsize:(idx + 1) * self.batchsize]
        step = -(-len(filenames) // max(1, chunks))
        args = self._read_args()
        return [pool.submit(_read_many_bytes, args, filenames[i:i + step]) for i in range(0, len(filenames), step)]

    def finish_batch(self, idx, futures):
        """-> (images uint8 or float32 [B, H, W, C], z or None, filenames); call in batch order (z uses numpy's global
        generator exactly like `__getitem__`)."""
        filenames = self.data[idx * self.batchsize:(idx + 1) * self.batchsize]
        parts = [f.result() for f in futures]
        if not all(p.dtype == np.uint8 for p in parts):
            parts = [_TO_UNIT[p] if p.dtype == np.uint8 else p for p in parts]
        images = np.concatenate(parts, axis=0)
        if self.phase == 'test':
            return images, None, filenames
        batch_z = np.random.normal(size=(self.batchsize, self.config['z_dim']))
        if self.num_classes is not None:
            classes = np.array([self.class_of(f) for f in filenames]).reshape((self.batchsize, 1))
            batch_z = np.concatenate((batch_z, classes), axis=1)
        return images, batch_z, filenames

    def __getitem__(self, idx):
        return self.load_batch(idx)


class DevicePrefetcher:
    """Keeps the GPU fed.  Worker processes decode, stretch and resize the files of the next `depth` batches and send
    back BYTES (the reference's last step, x / 127.5 - 1, is a 256-entry table); a producer thread collects the
    batches in order, draws z, stages the bytes in pinned host memory, copies them on a side stream and runs the
    table lookup there (`eg_u8_lut_f32`), so the training stream only waits on an event.  Iterating yields
    (images, z, filenames) with images / z as float32 device tensors of `ops`; a yielded batch stays valid until
    `depth` further batches have been taken.  On a CPU operator set (tests) the tensors are host tensors.  The random
    draws for z happen in batch order, so a seeded run sees the same z sequence as the sequential loader."""

    def __init__(self, dataset, ops, workers=8, depth=2, processes=True):
        self.dataset, self.ops, self.workers, self.depth = dataset, ops, max(1, workers), max(2, depth)
        self.processes = processes           # decode in worker processes (the GIL serialises PIL + numpy on threads)
        self._cuda = getattr(getattr(ops, "device", None), "type", "cpu") == "cuda"

    def __len__(self):
        return len(self.dataset)

    def _stage(self, slot, images, z):
        """images: uint8 (table lookup on the device) or float32; z: float64 or None -> device tensors + copy event"""
        import torch
        z32 = None if z is None else np.asarray(z, np.float32)
        if not self._cuda:
            im = _TO_UNIT[images] if images.dtype == np.uint8 else images
            out = [self.ops.from_numpy(np.asarray(im, np.float32))]
            if z32 is not None:
                out.append(self.ops.from_numpy(z32))
            return out, None
        if slot in self._copy_ev:
            self._copy_ev[slot].synchronize()            # the pinned buffers of this slot are free again
        if slot in self._done_ev:
            self._copy_stream.wait_event(self._done_ev.pop(slot))   # ... and the step that read its device buffers ran
        out = []
        with torch.cuda.stream(self._copy_stream):
            for j, a in enumerate([images] + ([z32] if z32 is not None else [])):
                key = (slot, j, a.dtype.str)
                tdt = torch.uint8 if a.dtype == np.uint8 else torch.float32
                if key not in self._bufs or tuple(self._bufs[key][0].shape) != a.shape:
                    self._bufs[key] = (torch.empty(a.shape, dtype=tdt).pin_memory(),
                                       torch.empty(a.shape, dtype=tdt, device=self.ops.device),
                                       torch.empty(a.shape, dtype=torch.float32, device=self.ops.device) if tdt == torch.uint8 else None)
                host, dev, f32 = self._bufs[key]
                host.numpy()[...] = a
                dev.copy_(host, non_blocking=True)
                if f32 is not None:
                    self.ops.u8_lut(dev, self._lut, f32, stream=self._copy_stream)
                    dev = f32
                out.append(dev)
            ev = torch.cuda.Event()
            ev.record(self._copy_stream)
            self._copy_ev[slot] = ev
        return out, ev

    def __iter__(self):
        import torch
        q = queue.Queue(maxsize=self.depth - 1)
        stop = threading.Event()
        if self._cuda:
            self._copy_stream = torch.cuda.Stream(device=self.ops.device)
            self._bufs, self._copy_ev, self._done_ev = {}, {}, {}
            self._lut = torch.from_numpy(_TO_UNIT.copy()).to(self.ops.device)
        nslots = self.depth + 1
        ds, nb = self.dataset, len(self.dataset)

        def produce():
            try:
                with pool:
                    pending = {}
                    for idx in range(nb):
                        if stop.is_set():
                            return
                        for k in range(idx, min(nb, idx + self.depth + 1)):       # decode ahead
                            if k not in pending:
                                pending[k] = ds.submit_batch(k, pool, self.workers)
                        images, z, filenames = ds.finish_batch(idx, pending.pop(idx))
                        staged, ev = self._stage(idx % nslots, images, z)
                        q.put((staged, ev, filenames, idx % nslots))
                q.put(None)
            except BaseException as e:          # surface decode errors in the consumer
                q.put(e)

        # The pool is created here, on the consumer's thread, before the producer thread exists: worker processes are
        # forked (like torch's DataLoader; they only run PIL + numpy and never touch CUDA), and a fork context does
        # not re-import the caller's __main__.
        if self.processes and self.workers > 1:
            import multiprocessing as mp
            from concurrent.futures import ProcessPoolExecutor
            pool = ProcessPoolExecutor(self.workers, mp_context=mp.get_context("fork"))
        else:
            pool = ThreadPoolExecutor(self.workers)
        t = threading.Thread(target=produce, daemon=True)
        t.start()
        last_slot = None
        try:
            while True:
                if last_slot is not None:
                    # The consumer is back for the next batch, so all work on the previous one has been enqueued.  Its
                    # "done" event must be published BEFORE the dequeue below: that dequeue is what lets the producer
                    # move on to staging a later batch into this very slot, and it only waits on events it can see.
                    done = torch.cuda.Event()
                    done.record(torch.cuda.current_stream(self.ops.device))
                    self._done_ev[last_slot] = done
                    last_slot = None
                item = q.get()
                if item is None:
                    return
                if isinstance(item, BaseException):
                    raise item
                staged, ev, filenames, slot = item
                if ev is not None:
                    torch.cuda.current_stream(self.ops.device).wait_event(ev)
                    last_slot = slot
                yield (*staged, filenames)
        finally:
            stop.set()
            while t.is_alive():
                try:
                    q.get_nowait()
                except queue.Empty:
                    t.join(0.05)
