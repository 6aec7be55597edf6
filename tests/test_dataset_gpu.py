"""N1 input pipeline: the CUDA preprocessing kernel and the device-resident loader against the reference's own image
arithmetic (dataset.py:37-73 restated on PIL + cv2 in oracle/preprocess.py) -- byte/integer work, so BIT-EXACT."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def rand_image(rng, H, W):
    """Smooth-ish random RGB content with hard edges (thin dark lines on white like an edge map on the left half)."""
    img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    img[:, : W // 2] = 255
    for _ in range(12):
        r, c = rng.integers(0, H), rng.integers(0, max(1, W // 2))
        img[r, : W // 2] = rng.integers(0, 80)
        img[:, c] = rng.integers(0, 80)
    return img


@pytest.mark.parametrize("H,W,domain,S", [
    (256, 512, "A", 64), (256, 512, "B", 64), (256, 512, "A", 512), (256, 512, "B", 512),      # edges2shoes pairs
    (218, 178, None, 64), (218, 178, None, 512),                                                # CelebA aligned crops
    (97, 131, None, 128), (600, 300, None, 64), (256, 600, "B", 64), (64, 64, None, 64), (33, 300, "A", 64),
])
def test_preprocess_kernel_bit_exact(H, W, domain, S):
    from discogan_modernized_b200.dataset import DeviceImageStore
    from oracle.preprocess import preprocess_cv2, preprocess_restated
    rng = np.random.default_rng(H * 1000 + W + S)
    imgs = [rand_image(rng, H, W) for _ in range(3)]
    store = DeviceImageStore(domain)
    for im in imgs:
        store.add(im)
    got = store.batch([2, 0, 1, 0], S).cpu().numpy()
    for j, i in enumerate([2, 0, 1, 0]):
        want = preprocess_cv2(imgs[i], domain, S)
        assert np.array_equal(preprocess_restated(imgs[i], domain, S), want)
        assert got[j].shape == want.shape and got[j].dtype == np.float32
        assert np.array_equal(got[j], want), (i, float(np.abs(got[j] - want).max()))


def test_ragged_batch_and_empty():
    """Images of different sizes in one launch; an empty index list is a no-op."""
    from discogan_modernized_b200.dataset import DeviceImageStore
    from oracle.preprocess import preprocess_cv2
    rng = np.random.default_rng(7)
    imgs = [rand_image(rng, h, w) for h, w in ((40, 50), (300, 200), (64, 64), (500, 700))]
    store = DeviceImageStore(None, chunk_bytes=1 << 16)          # forces several chunks
    for im in imgs:
        store.add(im)
    got = store.batch([0, 1, 2, 3], 64).cpu().numpy()
    for i, im in enumerate(imgs):
        assert np.array_equal(got[i], preprocess_cv2(im, None, 64))
    assert len(store.chunks) >= 2
    assert store.batch([], 64).shape == (0, 3, 64, 64)
    with pytest.raises(ValueError):
        DeviceImageStore("B").add(np.zeros((10, 200, 3), np.uint8))          # no right half to crop
    with pytest.raises(ValueError):
        store.add(np.zeros((10, 10), np.uint8))


@pytest.fixture(scope="module")
def pair_folder(tmp_path_factory):
    from PIL import Image
    root = tmp_path_factory.mktemp("edges2shoes")
    rng = np.random.default_rng(3)
    for i in range(23):
        Image.fromarray(rand_image(rng, 256, 512)).save(root / f"{i:03d}.png")      # lossless: decode is exact
    for i in range(4):
        Image.fromarray(rand_image(rng, 256, 512)).save(root / f"j{i}.jpg", quality=90)
    return root


def test_read_images_matches_reference(pair_folder):
    from discogan_modernized_b200 import dataset
    from oracle.preprocess import read_images_cv2
    files = dataset.list_images(pair_folder)
    assert len(files) == 27
    for domain in ("A", "B"):
        got = dataset.read_images(files, domain, 64)
        want = read_images_cv2(files, domain, 64)
        assert got.is_cuda and np.array_equal(got.cpu().numpy(), want)
    arr = dataset.read_images(files[:3], "B", 128, as_numpy=True)
    assert isinstance(arr, np.ndarray) and arr.shape == (3, 3, 128, 128)
    with pytest.raises(ValueError):
        dataset.read_images([str(pair_folder / "missing.png")], None, 64)


@pytest.mark.parametrize("resident", [True, False])
def test_dataset_batches(pair_folder, resident):
    """DiscoGANDataset pairing (index i -> A[i % len A], B[i % len B]) through the DistributedSampler index stream, in
    resident (HBM cache) and streaming mode, two ranks: identical to the reference loader's arithmetic."""
    from discogan_modernized_b200 import dataset
    from oracle.preprocess import read_images_cv2
    files = dataset.list_images(pair_folder)
    A_paths, B_paths = files[:20], files[5:]
    ds = dataset.DiscoGANDataset(A_paths, B_paths, "A", "B", 64, cache_bytes=None if resident else 0)
    assert len(ds) == 20 and ds.load() == resident
    a, b = ds[21]
    assert np.array_equal(a.cpu().numpy(), read_images_cv2([A_paths[1]], "A", 64)[0])
    assert np.array_equal(b.cpu().numpy(), read_images_cv2([B_paths[21]], "B", 64)[0])
    seen = []
    for rank in range(2):
        idx = dataset.sampler_indices(len(ds), rank, 2, epoch=3, seed=0)
        ref = torch.utils.data.distributed.DistributedSampler(range(len(ds)), num_replicas=2, rank=rank, shuffle=True, seed=0)
        ref.set_epoch(3)
        assert idx.tolist() == list(ref)
        got = [(A.clone(), B.clone()) for A, B in ds.batches(4, epoch=3, rank=rank, world=2)]
        assert [x[0].shape[0] for x in got] == [4, 4, 2]
        A_all, B_all = torch.cat([x[0] for x in got]), torch.cat([x[1] for x in got])
        wantA = read_images_cv2([A_paths[i % 20] for i in idx.tolist()], "A", 64)
        wantB = read_images_cv2([B_paths[i % 22] for i in idx.tolist()], "B", 64)
        assert np.array_equal(A_all.cpu().numpy(), wantA) and np.array_equal(B_all.cpu().numpy(), wantB)
        seen += idx.tolist()
    assert sorted(seen) == list(range(20))
    # image_translation.py style: independent shuffles, consecutive slices, drop the ragged tail
    ind = list(ds.batches(8, epoch=0, independent=True, drop_last=True))
    assert [x[0].shape[0] for x in ind] == [8, 8]


def test_loader_feeds_the_trainer(pair_folder):
    """The loader's ring buffers go straight into DiscoGANTrainer.step (graph replay copies them into its static inputs)."""
    from discogan_modernized_b200 import DiscoGANTrainer, dataset
    files = dataset.list_images(pair_folder)
    ds = dataset.DiscoGANDataset(files, files, "A", "B", 64)
    tr = DiscoGANTrainer(image_size=64, seed=1234, data_parallel=False)
    n = 0
    for epoch in range(3):
        for A, B in ds.batches(8, epoch=epoch, drop_last=True):
            tr.step(A, B)
            n += 1
    assert n == 9 and tr.iters == 9
    l = tr.losses()
    assert all(np.isfinite(v) for v in l.values()) and 0.0 < l["recon_loss_A"] < 1.0
    tr.close()
