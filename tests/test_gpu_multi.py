"""GPU tests (-m gpu) of one frame on several GPUs inside one process (wpt_scene_create with WPT_DEVICE_ALL):
the replacement of running the reference under MPI (mpi.hpp:42-289). Skipped on a box with one GPU."""
import numpy as np
import pytest

from wurblpt_b200 import abi
from util import Golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu(capi):
    if capi.device_count() == 0:
        pytest.fail("no CUDA device: the GPU tests must run on the B200 box (there is no CPU fallback)")
    if capi.device_count() < 2:
        pytest.skip("one frame on several GPUs needs at least two devices")
    return capi


@pytest.mark.parametrize("name,substreams", [("cornell", 1), ("cornell", 8), ("rtiow", 4), ("tof", 2), ("media", 4)])
def test_all_devices_render_the_frame_one_device_renders(gpu, name, substreams):
    """Every work item (pixel, sub-stream) is rendered by exactly one device with the same Prng stream, the partial
    sums of a pixel are added in a fixed order on the device that rendered it, and the frames are combined by adding
    zeros: the frame must equal the single-GPU frame BIT FOR BIT, whichever device drew which block of items."""
    g = Golden(name)
    one = gpu.Scene(g.desc, device=0).update_bvh(g.t0, g.t1)
    many = gpu.Scene(g.desc, device=abi.WPT_DEVICE_ALL).update_bvh(g.t0, g.t1)
    assert many.device_count() == gpu.device_count() >= 2
    sensor = g.sensor
    sensor.height, sensor.width = g.z["render"].shape[:2]
    spp = int(g.z["render_spp_sqrt"])
    f1, s1 = one.render(g.camera, sensor, g.params, spp, g.t0, g.t1, gpu.make_opts(substreams=substreams))
    for _ in range(2):      # twice: the block hand-out differs from run to run, the frame must not
        fn, sn = many.render(g.camera, sensor, g.params, spp, g.t0, g.t1, gpu.make_opts(substreams=substreams))
        assert sn.samples == s1.samples
        if g.desc.n_media == 0:
            assert sn.rays == s1.rays
        assert np.array_equal(fn.view(np.uint32), f1.view(np.uint32)), name
    one.close()
    many.close()


def test_devices_share_the_work(gpu):
    """A frame large enough to keep every device busy: all of them must have rendered part of it (the shared counter
    hands blocks to whoever asks), and queries / ground truth still work on the multi-device scene."""
    g = Golden("cornell")
    many = gpu.Scene(g.desc, device=abi.WPT_DEVICE_ALL).update_bvh(g.t0, g.t1)
    sensor = g.sensor
    sensor.height, sensor.width = 512, 512
    frame, stats = many.render(g.camera, sensor, g.params, 8, g.t0, g.t1)
    assert stats.samples == 512 * 512 * 64
    assert np.isfinite(frame).all() and frame.mean() > 0.01
    one = gpu.Scene(g.desc, device=0).update_bvh(g.t0, g.t1)
    ids, dist = many.primary_hits(g.camera, 96, 96, g.t0, g.params.min_hit_distance)
    ids1, dist1 = one.primary_hits(g.camera, 96, 96, g.t0, g.params.min_hit_distance)
    assert np.array_equal(ids, ids1) and np.array_equal(dist.view(np.uint32), dist1.view(np.uint32))
    one.close()
    many.close()
