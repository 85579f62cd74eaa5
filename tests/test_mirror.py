"""CPU tests of the drop-in boundary: the C++ API mirror (include/wurblpt) must describe scenes
exactly like the reference does, the reference's example applications must compile against it
unchanged, and the libtgd stand-in must round-trip what it writes."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = "/root/reference"


@pytest.fixture(scope="module")
def scenes():
    from wurblpt_b200 import scenes as s
    try:
        s.lib()
    except Exception as e:
        pytest.skip("libwpt_scenes.so not built: %s" % e)
    return s


@pytest.mark.parametrize("name,scale", [("cornell", 1.0), ("rtiow", 1.0), ("tof", 1.0), ("media", 1.0), ("mis", 1.0),
                                        ("mis_nohot", 1.0), ("furnace", 1.0), ("mixed", 1.0), ("synth", 0.004)])
def test_mirror_description_equals_golden(scenes, name, scale):
    """The mirror's flattening of scenes/configs.hpp is byte-identical to the description the oracle
    harness derived from the reference's own Scene object (stored in tests/golden)."""
    from wurblpt_b200 import scene_desc as sd
    from util import Golden
    g = Golden(name)
    built = scenes.BuiltScene(name, scale)
    mine = sd.desc_to_arrays(built.desc)
    assert sd.diff_descs(mine, g.host.arrays) == []
    v = scenes.View(name)
    assert bytes(v.params) == bytes(g.params)
    assert bytes(v.camera)[:120] == bytes(g.camera)[:120]      # everything but the key-frame pointer
    built.close()


def test_mirror_description_equals_live_reference_1m_triangles(scenes, ref):
    from wurblpt_b200 import scene_desc as sd
    a = sd.desc_to_arrays(scenes.BuiltScene("synth", 1.0).desc)
    b = sd.desc_to_arrays(ref.RefScene("synth", 1.0, 0).desc)
    assert sd.diff_descs(a, b) == []


def test_obj_import_equals_the_reference_importer(scenes, ref):
    """importIntoScene(): this repository's OBJ/MTL reader against the reference's tiny_obj_loader-based importer
    on a file with quads (split along the shorter diagonal), negative indices, missing normals and texture
    coordinates, a textured ModPhong material with texture options, an emitter (hot spot), a transparent material,
    an unknown material name and faces without material: the flattened scenes must be byte-identical."""
    from wurblpt_b200 import scene_desc as sd
    a = sd.desc_to_arrays(scenes.BuiltScene("imported", 1.0).desc)
    b = sd.desc_to_arrays(ref.RefScene("imported", 1.0, 0).desc)
    assert sd.diff_descs(a, b) == []
    assert a["meshes"].size > 0 and a["texels"].size == 48


def test_mirror_views_equal_reference_views(scenes, ref):
    for name in ["cornell", "rtiow", "synth", "synth360", "tof", "media", "furnace", "mis", "mixed", "imported"]:
        a, b = scenes.View(name), ref.View(name)
        assert bytes(a.sensor) == bytes(b.sensor), name
        assert bytes(a.params) == bytes(b.params), name
        assert bytes(a.camera)[:120] == bytes(b.camera)[:120], name
        assert (a.samples_sqrt, a.t0, a.t1) == (b.samples_sqrt, b.t0, b.t1), name


@pytest.mark.skipif(not os.path.exists(REFERENCE + "/wurblpt-cornellbox/wurblpt-cornellbox.cpp"),
                    reason="reference sources not present")
@pytest.mark.parametrize("app", ["wurblpt-animations", "wurblpt-envmap", "wurblpt-material-comparison",
                                 "wurblpt-normalmap", "wurblpt-rttnw", "wurblpt-stagelights", "wurblpt-rolling-marbles",
                                 "wurblpt-sponza", "wurblpt-bistro", "wurblpt-rungholt", "wurblpt-san-miguel",
                                 "wurblpt-tof-hcibox"])
def test_reference_app_source_is_accepted_by_the_mirror(app):
    """Apps that need image or OBJ files that are not available offline, or that render hours of 4K video, are not
    run here, but their unchanged sources must at least compile against the API mirror (with the eleven that are
    built and run, that is all 22 example apps of the reference)."""
    src = "%s/%s/%s.cpp" % (REFERENCE, app, app)
    r = subprocess.run(["/usr/bin/g++", "-std=c++20", "-O0", "-fsyntax-only", "-I" + os.path.join(ROOT, "include"), src],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]


@pytest.mark.skipif(not os.path.exists(REFERENCE + "/wurblpt-cornellbox/wurblpt-cornellbox.cpp"),
                    reason="reference sources not present")
@pytest.mark.parametrize("app", ["wurblpt-cornellbox", "wurblpt-rtiow", "wurblpt-furnace-test", "wurblpt-mis-test",
                                 "wurblpt-participating-media", "wurblpt-tof-example", "wurblpt-material-playground",
                                 "wurblpt-toomuch", "wurblpt-noise-textures"])
def test_reference_app_compiles_unchanged_against_the_mirror(app, tmp_path):
    """'existing example apps link unchanged': compile the reference's own app source against
    include/wurblpt + include/tgd and link it against libwurblpt_b200.so."""
    src = "%s/%s/%s.cpp" % (REFERENCE, app, app)
    out = str(tmp_path / app)
    lib = os.path.join(ROOT, "wurblpt_b200", "lib")
    cmd = ["/usr/bin/g++", "-std=c++20", "-O1", "-I" + os.path.join(ROOT, "include"), src, "-o", out,
           "-L" + lib, "-lwurblpt_b200", "-Wl,-rpath," + lib]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]


def test_mirror_refuses_to_render_without_a_device(tmp_path):
    """mcpt() through the mirror has no CPU fallback: without a GPU the app must stop with an error."""
    from wurblpt_b200 import capi
    if capi.device_count() > 0:
        pytest.skip("a GPU is present")
    src = tmp_path / "t.cpp"
    src.write_text('''
#include <wurblpt/wurblpt.hpp>
using namespace WurblPT;
int main() {
    Scene scene;
    scene.take(new Sphere(vec3(0.0f), 1.0f, scene.take(new MaterialLambertian(vec3(0.5f)))));
    SensorRGB sensor(8, 8);
    Camera camera(Optics(Projection(radians(40.0f), 1.0f)), Transformation(vec3(0.0f, 0.0f, 5.0f)));
    scene.updateBVH();
    mcpt(sensor, camera, scene, 1);
    return 0;
}
''')
    lib = os.path.join(ROOT, "wurblpt_b200", "lib")
    exe = str(tmp_path / "t")
    r = subprocess.run(["/usr/bin/g++", "-std=c++20", "-O1", "-I" + os.path.join(ROOT, "include"), str(src), "-o", exe,
                        "-L" + lib, "-lwurblpt_b200", "-Wl,-rpath," + lib], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode != 0
    assert "no CUDA device" in r.stderr or "CPU fallback" in r.stderr


def test_tgd_stand_in_round_trips(tmp_path):
    src = tmp_path / "io.cpp"
    src.write_text('''
#include <tgd/array.hpp>
#include <tgd/io.hpp>
#include <cstdio>
int main(int argc, char** argv) {
    TGD::Array<float> a({ size_t(5), size_t(3) }, 3);
    for (size_t i = 0; i < a.elementCount() * 3; i++) a.begin()[i] = 0.25f * i;
    a.globalTagList().set("K", "V");
    const char* names[] = { "a.pfm", "a.tgd" };
    for (const char* n : names) {
        std::string path = std::string(argv[1]) + "/" + n;
        if (TGD::save(a, path) != TGD::ErrorNone) return 1;
        TGD::Error e;
        TGD::ArrayContainer b = TGD::load(path, TGD::TagList(), &e);
        if (e != TGD::ErrorNone || b.dimension(0) != 5 || b.dimension(1) != 3 || b.componentCount() != 3) return 2;
        for (size_t i = 0; i < 45; i++) if (static_cast<const float*>(b.data())[i] != 0.25f * i) return 3;
    }
    TGD::Error e;
    if (TGD::load(std::string(argv[1]) + "/missing.png", TGD::TagList(), &e).elementCount() != 0 || e == TGD::ErrorNone) return 4;
    TGD::Array<float> c = a;    // shares storage, like libtgd containers
    c[{ 1, 1 }][2] = 42.0f;
    if (a[{ 1, 1 }][2] != 42.0f) return 5;
    return 0;
}
''')
    exe = str(tmp_path / "io")
    r = subprocess.run(["/usr/bin/g++", "-std=c++20", "-I" + os.path.join(ROOT, "include"), str(src), "-o", exe],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    assert subprocess.run([exe, str(tmp_path)]).returncode == 0


def test_user_defined_animations_are_sampled_into_key_frames(tmp_path):
    """an Animation that is not an AnimationKeyframes (as in the reference's
    material-playground / animations / toomuch apps) is sampled at the times a description is made for:
    one exact key frame for t0 == t1, 17 across an exposure interval; the camera likewise."""
    src = tmp_path / "bake.cpp"
    src.write_text('''
#include <cstdio>
#include <wurblpt/wurblpt.hpp>
using namespace WurblPT;
class Orbit : public Animation {
public:
    virtual Transformation at(float t) const {
        Transformation T;
        T.translate(vec3(cos(t), sin(t), -1.0f));
        T.rotate(toQuat(t, vec3(0.0f, 1.0f, 0.0f)));
        return T;
    }
};
int main() {
    Scene scene;
    Material* m = scene.take(new MaterialLambertian(vec3(0.5f)));
    int orbit = scene.take(new Orbit);
    scene.take(new MeshInstance(scene.take(generateCube()), m, Transformation(), orbit));
    Orbit reference;
    for (int pass = 0; pass < 2; pass++) {
        float t0 = 0.5f, t1 = (pass == 0 ? 0.5f : 0.75f);
        wpt_scene_desc d = scene.describe(t0, t1);
        if (d.n_animations != 1 || !scene.hasBakedAnimations()) return 2;
        unsigned int n = d.animations[0].keyframe_count;
        std::printf("pass %d keyframes %u\\n", pass, n);
        for (unsigned int i = 0; i < n; i++) {
            const wpt_keyframe& k = d.keyframes[d.animations[0].first_keyframe + i];
            Transformation T = reference.at(k.t);
            if (k.translation[0] != T.translation.x() || k.translation[1] != T.translation.y()
                    || k.rotation[1] != T.rotation.y || k.rotation[3] != T.rotation.w) return 3;
            if (i == 0 && k.t != t0) return 4;
            if (i == n - 1 && k.t != t1) return 5;
        }
    }
    Camera camera(Optics(Projection(radians(40.0f), 1.0f)), new Orbit);
    wpt_camera_desc c;
    std::vector<wpt_keyframe> keys;
    if (camera.flatten(c, keys)) return 6;                                   // no times given: cannot be described
    if (!camera.flatten(c, keys, detail::bakeTimesFor(1.0f, 1.0f)) || c.n_keyframes != 1 || keys[0].t != 1.0f) return 7;
    std::printf("ok\\n");
    return 0;
}
''')
    exe = str(tmp_path / "bake")
    lib = os.path.join(ROOT, "wurblpt_b200", "lib")
    r = subprocess.run(["/usr/bin/g++", "-std=c++20", "-O1", "-I" + os.path.join(ROOT, "include"), str(src), "-o", exe,
                        "-L" + lib, "-lwurblpt_b200", "-Wl,-rpath," + lib], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr[-500:])
    assert "pass 0 keyframes 1" in r.stdout and "pass 1 keyframes 17" in r.stdout and "ok" in r.stdout


NOISE_PROBE = r'''
#include <cstdio>
#include <cstring>
#include <wurblpt/wurblpt.hpp>
using namespace WurblPT;
static void show(const vec4& v) {
    for (int c = 0; c < 4; c++) { float f = v[c]; unsigned int u; std::memcpy(&u, &f, 4); std::printf("%08x ", u); }
    std::printf("\n");
}
int main() {
    Prng prng(4711);
    TextureValueNoise value(7, 5, prng);
    TextureGradientNoise gradient(6, 9, prng);
    TextureWorleyNoise worley(13, prng);
    TexturePerlinNoise perlin(prng);
    Prng where(99);
    for (int i = 0; i < 200; i++) {
        vec2 tc = vec2(where.in01() * 3.0f, where.in01() * 2.0f);
        show(value.value(tc, 0.0f));
        show(gradient.value(tc, 0.0f));
        show(worley.value(tc, 0.0f));
        show(perlin.value(tc, 0.0f));
        show(perlin.turbulence(tc, 0.0f));
    }
    return 0;
}
'''


@pytest.mark.skipif(not os.path.exists(REFERENCE + "/libwurblpt/texture_noise.hpp"), reason="reference sources not present")
def test_noise_textures_equal_the_reference_bit_for_bit(tmp_path):
    """The same program compiled against the reference's headers and against the API mirror: the four noise
    textures (built from the same Prng) must return identical values, bit for bit."""
    src = tmp_path / "noise.cpp"
    src.write_text(NOISE_PROBE)
    outputs = []
    shim = os.path.join(ROOT, "oracle", "_ref", "inc")      # <wurblpt/...> -> the reference's libwurblpt
    for name, includes, link in (
            ("ref", ["-I" + shim, "-I" + os.path.join(ROOT, "include")], []),
            ("mirror", ["-I" + os.path.join(ROOT, "include")],
             ["-L" + os.path.join(ROOT, "wurblpt_b200", "lib"), "-lwurblpt_b200",
              "-Wl,-rpath," + os.path.join(ROOT, "wurblpt_b200", "lib")])):
        if name == "ref" and not os.path.isdir(shim):
            pytest.skip("oracle/_ref/inc not prepared (make -C oracle ref)")
        exe = str(tmp_path / name)
        extra = [REFERENCE + "/libwurblpt/powitacq.cpp"] if name == "ref" else []
        r = subprocess.run(["/usr/bin/g++", "-std=c++20", "-O1", "-ffp-contract=off", "-fopenmp"] + includes
                           + [str(src)] + extra + ["-o", exe] + link, capture_output=True, text=True)
        assert r.returncode == 0, (name, r.stderr[-2000:])
        outputs.append(subprocess.run([exe], capture_output=True, text=True).stdout)
    assert len(outputs[0].splitlines()) == 1000
    assert outputs[0] == outputs[1]


def test_mpi_coordinator_interface_without_mpi(tmp_path):
    """MPICoordinator (reference mpi.hpp:152-289) keeps the members applications use; there is one process, the world
    size is the number of GPUs that render a frame, the block interface hands out the image as one block, and the
    program refuses to run as one of several MPI ranks (every rank would render everything on the same GPUs)."""
    src = tmp_path / "t.cpp"
    src.write_text('''
#include <cstdio>
#include <wurblpt/wurblpt.hpp>
using namespace WurblPT;
int main() {
    MPICoordinator mpi;
    unsigned int start = 7, size = 7;
    float pixels[12];
    mpi.init(2, 2, pixels, 3);
    mpi.getBlock(&start, &size);
    std::printf("%d %d %u %u %u", mpi.mainProcess() ? 1 : 0, mpi.worldSize(), mpi.blockSize(), start, size);
    mpi.submitBlock(start, size);
    mpi.getBlock(&start, &size);
    std::printf(" %u %s\\n", size, mpi.processId());
    mpi.finish();
    return 0;
}
''')
    lib = os.path.join(ROOT, "wurblpt_b200", "lib")
    exe = str(tmp_path / "t")
    r = subprocess.run(["/usr/bin/g++", "-std=c++20", "-O1", "-I" + os.path.join(ROOT, "include"), str(src), "-o", exe,
                        "-L" + lib, "-lwurblpt_b200", "-Wl,-rpath," + lib], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    env = {k: v for k, v in os.environ.items() if not k.startswith(("OMPI_", "PMI_", "PMIX_", "MV2_"))}
    r = subprocess.run([exe], capture_output=True, text=True, env=dict(env, WURBLPT_DEVICE="0"))
    assert r.returncode == 0, r.stderr
    main, world, block, start, size, size2, *name = r.stdout.split()
    assert (main, world, block, start, size, size2) == ("1", "1", "4096", "0", "4", "0")   # one device chosen explicitly
    assert " ".join(name) == "process 0"
    from wurblpt_b200 import capi
    r = subprocess.run([exe], capture_output=True, text=True, env=env)
    assert r.returncode == 0 and int(r.stdout.split()[1]) == max(1, capi.device_count())     # all GPUs of the box
    r = subprocess.run([exe], capture_output=True, text=True, env=dict(env, OMPI_COMM_WORLD_SIZE="4", OMPI_COMM_WORLD_RANK="1"))
    assert r.returncode != 0 and "MPI ranks" in r.stderr
