"""Input pipeline for the B200 train step: host mirror of the reference's ``dataset.py`` image path
(``read_images`` :37-73, ``shuffle_data`` :24-35, ``DiscoGANDataset`` :194-261) and of the loaders built on it
(``image_translation.py:309-333``, ``distributed_image_translation.py:182-226``).

The reference decodes, crops, dilates, resizes and normalises every image on the host for every batch (PIL + cv2 per
image, pageable ``FloatTensor(...).to(device)``): at 2 ms per step that is the whole wall clock.  Here the only host work
is the JPEG/PNG decode, done ONCE per file by a thread pool; the decoded uint8 images live in HBM (a 200 k-image CelebA
is 23 GB of the 180 GB), and each batch is produced by one launch of ``dg_preprocess_u8`` -- crop, 3x3 edge thickening,
bilinear resize, /255, HWC->CHW, bit-exact with the reference's cv2 arithmetic -- straight into a ring of fp32
``[B,3,S,S]`` device buffers that the trainer consumes.  Datasets that do not fit the cache budget stream instead:
a producer thread decodes the next batches into pinned staging while the current step runs.

There is no CPU fallback for the arithmetic: the module needs the CUDA library.
"""
import os
import queue
import threading
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np
import torch

from . import ops

_ALIGN = 256


def decode_rgb(path):
    """File -> uint8 [H,W,3] (PIL, ``Image.open(fn).convert('RGB')`` as dataset.py:44,240)."""
    from PIL import Image
    with Image.open(path) as im:
        return np.asarray(im.convert("RGB"))


def domain_crop(domain, width):
    """(x0, crop_width, mode) for the reference's domain types (dataset.py:52-60): 'A' = left 256 columns of a
    side-by-side pair + edge thickening, 'B' = the columns from 256 on, None = the whole image."""
    if domain == "A":
        return 0, min(256, width), 1
    if domain == "B":
        if width <= 256:
            raise ValueError(f"domain 'B' needs an image wider than 256 pixels, got {width}")
        return 256, width - 256, 0
    if domain is None:
        return 0, width, 0
    raise ValueError(f"unknown domain type {domain!r}")


def task_domains(task_name):
    """Domain types by task, as image_translation.py:243-251 / distributed_image_translation.py:184-190."""
    if task_name.startswith("edges2"):
        return "A", "B"
    if task_name in ("handbags2shoes", "shoes2handbags"):
        return "B", "B"
    return None, None


def list_images(folder):
    """sorted *.jpg + *.png of a folder (dataset.py:131,181-182)."""
    folder = Path(folder)
    files = sorted(str(f) for ext in ("*.jpg", "*.jpeg", "*.png") for f in folder.glob(ext))
    if not files:
        raise FileNotFoundError(f"no .jpg/.png images under {folder}")
    return files


class DeviceImageStore:
    """Decoded uint8 images of one domain, resident in device memory, plus the ``dg_preprocess_u8`` table row of each."""

    def __init__(self, domain=None, device="cuda", chunk_bytes=256 << 20):
        self.domain, self.device = domain, torch.device(device)
        self.chunk_bytes = chunk_bytes
        self.chunks, self.rows = [], []
        self._stage = None          # (pinned uint8 buffer, fill, [(row index, offset)])
        self._table = None
        self.bytes = 0

    def __len__(self):
        return len(self.rows)

    def add(self, image_u8):
        """Append one decoded image (uint8 [H,W,3]); returns its index.  Data is staged in pinned memory and uploaded
        chunk-wise by ``flush``."""
        if image_u8.dtype != np.uint8 or image_u8.ndim != 3 or image_u8.shape[2] != 3:
            raise ValueError(f"expected a uint8 [H,W,3] image, got {image_u8.dtype} {image_u8.shape}")
        H, W, _ = image_u8.shape
        x0, cw, mode = domain_crop(self.domain, W)
        n = H * W * 3
        need = (n + _ALIGN - 1) // _ALIGN * _ALIGN
        if self._stage is not None and self._stage[1] + need > self._stage[0].numel():
            self.flush()
        if self._stage is None:
            self._stage = [torch.empty(max(self.chunk_bytes, need), dtype=torch.uint8).pin_memory(), 0, []]
        buf, fill, pend = self._stage
        buf[fill:fill + n].copy_(torch.from_numpy(np.ascontiguousarray(image_u8).reshape(-1)))
        pend.append((len(self.rows), fill))
        self._stage[1] = fill + need
        self.rows.append([0, H, W, x0, cw, mode, 0, 0])
        self.bytes += n
        self._table = None
        return len(self.rows) - 1

    def flush(self):
        if self._stage is None:
            return
        buf, fill, pend = self._stage
        dev = torch.empty(fill, dtype=torch.uint8, device=self.device)
        dev.copy_(buf[:fill], non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()      # the pinned buffer is reused / dropped right away
        self.chunks.append(dev)
        for row, off in pend:
            self.rows[row][0] = dev.data_ptr() + off
        self._stage, self._table = None, None

    def add_files(self, paths, workers=None):
        workers = workers or min(32, os.cpu_count() or 4)
        with ThreadPoolExecutor(workers) as pool:
            for img in pool.map(decode_rgb, paths):          # decode in parallel (PIL releases the GIL), append in order
                self.add(img)
        self.flush()
        return self

    def table(self):
        """Device int64 [N,8] with one row per image."""
        if self._table is None:
            self.flush()
            self._table = torch.tensor(self.rows, dtype=torch.int64, device=self.device).view(-1, 8)
        return self._table

    def batch(self, indices, image_size, out=None):
        """Preprocess the images with these indices (device or host int64 tensor / list) -> fp32 [n,3,S,S] on the device."""
        idx = torch.as_tensor(indices, dtype=torch.int64, device=self.device)
        return ops.preprocess_u8(self.table().index_select(0, idx), image_size, out=out)


def read_images(filenames, domain=None, image_size=64, device="cuda", workers=None, as_numpy=False):
    """``read_images`` (dataset.py:37-73): list of files -> [N,3,S,S] float32 in [0,1].  Decoding happens on host
    threads, everything else in one kernel launch.  Returns a CUDA tensor (``as_numpy=True``: the reference's numpy
    array).  Files that fail to decode are skipped with a message, as in the reference; no valid image raises."""
    store = DeviceImageStore(domain, device)
    workers = workers or min(32, os.cpu_count() or 4)

    def safe(fn):
        try:
            return decode_rgb(fn)
        except Exception as e:  # noqa: BLE001 -- dataset.py:45-47 prints and skips
            print(f"image load failed: {fn}: {e}")
            return None
    with ThreadPoolExecutor(workers) as pool:
        for img in pool.map(safe, list(filenames)):
            if img is not None:
                store.add(img)
    if len(store) == 0:
        raise ValueError("no valid images")
    out = store.batch(list(range(len(store))), image_size)
    return out.cpu().numpy() if as_numpy else out


def shuffle_data(da, db, rng=None):
    """``shuffle_data`` (dataset.py:24-35): independent permutations of the two domains."""
    rng = np.random if rng is None else rng
    a_idx, b_idx = np.arange(len(da)), np.arange(len(db))
    rng.shuffle(a_idx)
    rng.shuffle(b_idx)
    return np.array(da)[a_idx], np.array(db)[b_idx]


def sampler_indices(length, rank=0, world=1, epoch=0, seed=0, shuffle=True):
    """The index stream of ``torch.utils.data.DistributedSampler(dataset, world, rank, shuffle)`` after
    ``set_epoch(epoch)`` (distributed_image_translation.py:203-208,438-440): seed+epoch permutation, padded by wrapping to
    a multiple of the world size, rank-strided."""
    if shuffle:
        g = torch.Generator().manual_seed(seed + epoch)
        idx = torch.randperm(length, generator=g)
    else:
        idx = torch.arange(length)
    total = -(-length // world) * world
    if total > length:
        idx = torch.cat([idx, idx[:total - length]])
    return idx[rank:total:world]


class DiscoGANDataset:
    """``DiscoGANDataset`` (dataset.py:194-261): index i pairs ``A[i % len(A)]`` with ``B[i % len(B)]``; length is the
    shorter domain.  ``cache_bytes`` bounds the HBM spent on decoded images (default: half of the free device memory);
    larger datasets stream from the host."""

    def __init__(self, domain_A_paths, domain_B_paths, domain_A_type=None, domain_B_type=None, image_size=64,
                 transform=None, device="cuda", cache_bytes=None, workers=None):
        self.domain_A_paths, self.domain_B_paths = list(domain_A_paths), list(domain_B_paths)
        self.domain_A_type, self.domain_B_type = domain_A_type, domain_B_type
        self.image_size, self.transform = image_size, transform
        self.length = min(len(self.domain_A_paths), len(self.domain_B_paths))
        self.device = torch.device(device)
        self.workers = workers or min(32, os.cpu_count() or 4)
        if cache_bytes is None:
            free, _ = torch.cuda.mem_get_info(self.device)
            cache_bytes = free // 2
        self.cache_bytes = cache_bytes
        self._stores = None

    def __len__(self):
        return self.length

    # -- resident mode --------------------------------------------------------------------------------
    def _estimate_bytes(self):
        probe = [decode_rgb(p) for p in (self.domain_A_paths[0], self.domain_B_paths[0])]
        return probe[0].size * len(self.domain_A_paths) + probe[1].size * len(self.domain_B_paths)

    def load(self):
        """Decode every file once into the device store (resident mode).  Returns False if the estimate exceeds
        ``cache_bytes`` (callers then stream)."""
        if self._stores is not None:
            return True
        if self._estimate_bytes() > self.cache_bytes:
            return False
        self._stores = (DeviceImageStore(self.domain_A_type, self.device).add_files(self.domain_A_paths, self.workers),
                        DeviceImageStore(self.domain_B_type, self.device).add_files(self.domain_B_paths, self.workers))
        return True

    def __getitem__(self, index):
        """(a, b) fp32 [3,S,S] device tensors for pair ``index`` (dataset.py:215-236)."""
        ia, ib = index % len(self.domain_A_paths), index % len(self.domain_B_paths)
        if self.load():
            a = self._stores[0].batch([ia], self.image_size)[0]
            b = self._stores[1].batch([ib], self.image_size)[0]
        else:
            a = read_images([self.domain_A_paths[ia]], self.domain_A_type, self.image_size, self.device)[0]
            b = read_images([self.domain_B_paths[ib]], self.domain_B_type, self.image_size, self.device)[0]
        if self.transform:
            a, b = self.transform(a), self.transform(b)
        return a, b

    # -- batch streams --------------------------------------------------------------------------------
    def batches(self, batch_size, epoch=0, rank=0, world=1, shuffle=True, seed=0, independent=False, drop_last=False,
                ring=3, prefetch=3):
        """Yield (A, B) fp32 ``[b,3,S,S]`` device batches of one epoch.
        independent=False: the distributed loader -- DistributedSampler index stream, pair i = (A[i], B[i]);
        independent=True: ``image_translation.py:309-319`` -- the two domains shuffled independently (shuffle_data) and
        cut into consecutive slices.  Output buffers come from a ring of ``ring`` per domain: a batch stays valid until
        ``ring - 1`` further batches have been drawn."""
        S = self.image_size
        if independent:
            g = torch.Generator().manual_seed(seed + epoch)
            ia = torch.randperm(len(self.domain_A_paths), generator=g) if shuffle else torch.arange(len(self.domain_A_paths))
            ib = torch.randperm(len(self.domain_B_paths), generator=g) if shuffle else torch.arange(len(self.domain_B_paths))
            n = min(len(ia), len(ib)) // world
            ia, ib = ia[rank * n:(rank + 1) * n], ib[rank * n:(rank + 1) * n]
        else:
            idx = sampler_indices(self.length, rank, world, epoch, seed, shuffle)
            ia, ib = idx % len(self.domain_A_paths), idx % len(self.domain_B_paths)
        n = len(ia)
        starts = list(range(0, n - (n % batch_size if drop_last else 0), batch_size))
        bufs = [(torch.empty(batch_size, 3, S, S, device=self.device), torch.empty(batch_size, 3, S, S, device=self.device))
                for _ in range(ring)]
        if self.load():
            ia_d, ib_d = ia.to(self.device), ib.to(self.device)
            ta, tb = self._stores[0].table(), self._stores[1].table()
            for k, s in enumerate(starts):
                b = min(batch_size, n - s)
                oa, ob = bufs[k % ring]
                A = ops.preprocess_u8(ta.index_select(0, ia_d[s:s + b]), S, out=oa[:b])
                B = ops.preprocess_u8(tb.index_select(0, ib_d[s:s + b]), S, out=ob[:b])
                yield A, B
            return
        yield from self._stream(ia.tolist(), ib.tolist(), starts, batch_size, bufs, prefetch)

    def _stream(self, ia, ib, starts, batch_size, bufs, prefetch):
        """Streaming mode: a producer thread decodes the next batches on the pool into pinned staging; the consumer
        uploads on a copy stream and launches the preprocessing kernel behind the copy."""
        S, n, ring = self.image_size, len(ia), len(bufs)
        q = queue.Queue(maxsize=prefetch)
        pool = ThreadPoolExecutor(self.workers)

        def produce():
            try:
                for s in starts:
                    b = min(batch_size, n - s)
                    imgs_a = list(pool.map(decode_rgb, [self.domain_A_paths[i] for i in ia[s:s + b]]))
                    imgs_b = list(pool.map(decode_rgb, [self.domain_B_paths[i] for i in ib[s:s + b]]))
                    q.put((imgs_a, imgs_b))
                q.put(None)
            except Exception as e:  # noqa: BLE001 -- surfaced in the consumer
                q.put(e)

        threading.Thread(target=produce, daemon=True).start()
        copy_stream = torch.cuda.Stream(self.device)
        keep = []                                              # stores of the last `ring` batches stay alive
        k = 0
        while True:
            item = q.get()
            if item is None:
                break
            if isinstance(item, Exception):
                raise item
            oa, ob = bufs[k % ring]
            with torch.cuda.stream(copy_stream):
                stores = []
                outs = []
                for imgs, dom, o in ((item[0], self.domain_A_type, oa), (item[1], self.domain_B_type, ob)):
                    st = DeviceImageStore(dom, self.device, chunk_bytes=sum(i.size for i in imgs) + _ALIGN * len(imgs))
                    for im in imgs:
                        st.add(im)
                    outs.append(st.batch(list(range(len(st))), S, out=o[:len(imgs)]))
                    stores.append(st)
                done = torch.cuda.Event()
                done.record(copy_stream)
            torch.cuda.current_stream(self.device).wait_event(done)
            keep.append(stores)
            if len(keep) > ring:
                keep.pop(0)
            k += 1
            yield outs[0], outs[1]
        pool.shutdown(wait=False)
