"""The stand-alone host layer (host/rtb_scene.hpp via librtb200_host.so) against the reference's
own loader: the flattened scene — camera matrices, every triangle in BVH order, every BVH node,
materials, textures, lights — must be byte-identical, because hit-ID parity starts there."""
import os

import numpy as np
import pytest

from conftest import ref_scene


@pytest.fixture(scope="module")
def host():
    from raytracingrenderer_b200 import host_api
    host_api.lib()
    return host_api


@pytest.mark.parametrize("name", ["cornell-box", "MaterialsScene", "materialball", "coffee", "MaterialsScene_env",
                                  "materialball_glass", "materialball_layered", "materialball_conductor",
                                  "bathroom"])   # bathroom: baseline + progressive JPEG textures
def test_loader_and_builder_equal_the_reference(host, name, tmp_path):
    rs = ref_scene(name)
    want = rs.flatten(str(tmp_path / "ref.rtbs"))
    got = host.load_scene(rs.dir)
    assert got.camera.tobytes() == want.camera.tobytes()
    for k in ("ref_nodes", "tri_isect", "tri_shade", "materials", "textures", "lights", "texels"):
        a, b = getattr(got, k), getattr(want, k)
        assert len(a) == len(b), k
        assert a.tobytes() == b.tobytes(), k
    assert (got.background_type, got.background_tex) == (want.background_type, want.background_tex)
    assert got.background_colour.tobytes() == want.background_colour.tobytes()


def test_undecodable_textures_are_an_error_not_a_silent_default(host, tmp_path):
    """A texture file that exists but cannot be decoded (here: a JPEG cut off inside its header)
    must raise; only a MISSING file becomes the reference's 1x1 white default."""
    import shutil
    rs = ref_scene("bathroom")
    d = tmp_path / "scene"
    shutil.copytree(rs.dir, d, symlinks=False)
    jpgs = sorted(p for p in os.listdir(d) if p.lower().endswith((".jpg", ".jpeg")))
    assert jpgs
    victim = d / jpgs[0]
    data = open(victim, "rb").read()
    os.remove(victim)
    open(victim, "wb").write(data[:64])
    with pytest.raises(RuntimeError) as e:
        host.load_scene(str(d))
    assert "cannot decode" in str(e.value)


def test_soup_scene_is_well_formed(host, oracle_mod):
    from raytracingrenderer_b200 import abi
    s, secs = host.build_soup(4096, 160, 90)
    assert 4000 < s.n_tris <= 4096 and len(s.lights) == 1 and s.lights["type"][0] == abi.LIGHT_BACKGROUND
    leaves = s.ref_nodes[s.ref_nodes["a"] < 0]
    assert leaves["b"].max() <= 2 and int(leaves["b"].sum()) == s.n_tris          # MAXNODE_TRIANGLES
    assert sorted((~leaves["a"]).tolist()) == sorted(np.cumsum(np.r_[0, leaves["b"][np.argsort(~leaves["a"])]])[:-1].tolist())
    o = oracle_mod.Oracle(s, max_depth=0)
    ids, t = o.primary_hits()
    assert 0.3 < (ids != abi.MISS_ID).mean() < 1.0
    film, st = o.render(1)
    assert np.isfinite(film).all() and st["closest_rays"] <= 2 * st["samples"]     # primary + one bounce


def test_camera_matches_the_reference_flattened_camera(host):
    rs = ref_scene("cornell-box")
    import json
    sj = json.load(open(os.path.join(rs.dir, "scene.json")))
    v = lambda k: [float(x) for x in sj[k].split()]  # noqa: E731
    cam = host.camera(v("from"), v("to"), v("up"), float(sj["fov"]), int(sj["width"]), int(sj["height"]))
    want = rs.flatten("/tmp/_cam.rtbs").camera
    assert cam.tobytes() == want.tobytes()


def test_parallel_build_equals_serial_build(host, tmp_path, monkeypatch):
    """The reference-order builder forks threads for disjoint sub-ranges above 200 k triangles; the
    tree and triangle order must not depend on that."""
    a, _ = host.build_soup(1 << 18, 64, 36)
    monkeypatch.setenv("RTB_HOST_BUILD_SERIAL", "1")
    b, _ = host.build_soup(1 << 18, 64, 36)
    assert a.ref_nodes.tobytes() == b.ref_nodes.tobytes() and a.tri_isect.tobytes() == b.tri_isect.tobytes()


def test_parallel_sort_reproduces_std_sort_ties_included(host):
    """Triangle IDs are positions after the reference builder's std::sort by centroid, and equal
    centroids are common, so the builder's concurrent sort must give std::sort's exact permutation."""
    rng = np.random.default_rng(5)
    n = 300000
    cases = {
        "random": rng.random(n, dtype=np.float32),
        "heavy ties": np.floor(rng.random(n) * 97).astype(np.float32),
        "all equal": np.zeros(n, np.float32),
        "sorted": np.arange(n, dtype=np.float32),
        "reversed": np.arange(n, dtype=np.float32)[::-1],
        "sawtooth": (np.arange(n) % 1000).astype(np.float32),
        "organ pipe": np.minimum(np.arange(n), n - np.arange(n)).astype(np.float32),
        "two values": (rng.random(n) < 0.5).astype(np.float32),
        "tiny": rng.random(11, dtype=np.float32),
        "17": np.floor(rng.random(17) * 3).astype(np.float32),
    }
    for name, k in cases.items():
        for par in (0, 1, 4):
            assert host.sort_selftest(k, par), (name, par)


def test_image_decoders_against_the_stb_image_golden(host):
    """PNG and JPEG (baseline / progressive, 4:4:4 / 4:2:2 / 4:2:0, odd sizes, restart intervals, optimised
    tables, grey) decode to exactly the bytes the reference's stb_image produces (golden written by
    tests/golden/make_golden.py: images()), and to stb_image's live output when oracle/_ref is present."""
    import hashlib, json
    d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "images")
    gold = json.load(open(os.path.join(d, "stb_image_golden.json")))
    assert len(gold) >= 14
    from oracle import ref
    for name, g in gold.items():
        a = host.decode_image(os.path.join(d, name))
        assert list(a.shape) == g["shape"], name
        assert hashlib.sha256(a.tobytes()).hexdigest() == g["sha256"], name
        if ref.available():
            assert np.array_equal(a, ref.decode_image(os.path.join(d, name))), name
    # truncated and corrupt files are errors, not crashes
    data = open(os.path.join(d, "prog_420_q80.jpg"), "rb").read()
    for cut in (2, 20, 200):
        p = os.path.join(d, "..", "..", "..", "gpurun_out")
        os.makedirs(p, exist_ok=True)
        f = os.path.join(p, "_cut.jpg")
        open(f, "wb").write(data[:cut])
        try:
            host.decode_image(f)
        except RuntimeError:
            pass
    os.remove(f)


def test_jpeg_decoder_on_random_images_against_stb_image(host, tmp_path):
    """Beyond the committed fixtures: 60 random images (size 1..97, smooth / noisy / flat content) written by
    Pillow with random quality, chroma subsampling, progressive / optimised / restart settings, decoded by the
    product and by the reference's stb_image (needs oracle/_ref and Pillow: skipped otherwise)."""
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(2026)
    for i in range(60):
        w, h = int(rng.integers(1, 98)), int(rng.integers(1, 98))
        kind = i % 3
        if kind == 0:
            img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        elif kind == 1:
            y, x = np.mgrid[0:h, 0:w]
            img = np.stack([(x * 5 + y * 3) % 256, (x * y) % 256, 255 - (x + y) % 256], -1).astype(np.uint8)
        else:
            img = np.full((h, w, 3), rng.integers(0, 256, 3), np.uint8)
        kw = dict(quality=int(rng.integers(5, 100)), subsampling=int(rng.integers(0, 3)))
        if rng.random() < 0.5:
            kw["progressive"] = True
        if rng.random() < 0.3:
            kw["optimize"] = True
        if rng.random() < 0.3:
            kw["restart_marker_blocks"] = int(rng.integers(1, 9))
        grey = rng.random() < 0.15
        f = str(tmp_path / ("r%02d.jpg" % i))
        (Image.fromarray(img[:, :, 0]) if grey else Image.fromarray(img)).save(f, "JPEG", **kw)
        a, b = host.decode_image(f), ref.decode_image(f)
        assert a.shape == b.shape and np.array_equal(a, b), (i, w, h, kw, grey)


def _write_png(path, w, h, ctype, depth, samples, interlace=False, palette=None, trns=None, rng=None):
    """Minimal PNG encoder for tests: `samples` is uint16 [h, w, n] (n = channels of the colour type, values <
    2^depth), every scanline gets a random filter type, Adam7 optional."""
    import struct, zlib
    n = samples.shape[2]

    def pack(rows):                      # rows: [y, x, n] -> bytes per scanline
        out = []
        for r in rows:
            v = r.reshape(-1)
            if depth == 16:
                b = v.astype(">u2").tobytes()
            elif depth == 8:
                b = v.astype(np.uint8).tobytes()
            else:
                bits = np.zeros(len(v) * depth, np.uint8)
                for k in range(depth):
                    bits[k::depth] = (v >> (depth - 1 - k)) & 1
                b = np.packbits(bits).tobytes()
            out.append(b)
        return out

    def filt(lines, bpp):
        res, prev = b"", None
        for line in lines:
            cur = np.frombuffer(line, np.uint8).astype(np.int32)
            up = np.zeros_like(cur) if prev is None else prev
            a = np.concatenate([np.zeros(bpp, np.int32), cur[:-bpp]]) if len(cur) > bpp else np.zeros_like(cur)
            c = np.concatenate([np.zeros(bpp, np.int32), up[:-bpp]]) if len(cur) > bpp else np.zeros_like(cur)
            if len(cur) <= bpp:
                a = np.zeros_like(cur); c = np.zeros_like(cur)
            ft = int(rng.integers(0, 5)) if rng is not None else 0
            if ft == 0: o = cur
            elif ft == 1: o = cur - a
            elif ft == 2: o = cur - up
            elif ft == 3: o = cur - ((a + up) >> 1)
            else:
                pp = a + up - c
                pa, pb, pc = np.abs(pp - a), np.abs(pp - up), np.abs(pp - c)
                pred = np.where((pa <= pb) & (pa <= pc), a, np.where(pb <= pc, up, c))
                o = cur - pred
            res += bytes([ft]) + (o & 255).astype(np.uint8).tobytes()
            prev = cur
        return res
    bpp = max(1, n * depth // 8)
    if not interlace:
        raw = filt(pack(samples), bpp)
    else:
        raw = b""
        for xo, yo, xs, ys in ((0, 0, 8, 8), (4, 0, 8, 8), (0, 4, 4, 8), (2, 0, 4, 4), (0, 2, 2, 4), (1, 0, 2, 2), (0, 1, 1, 2)):
            sub = samples[yo::ys, xo::xs]
            if sub.shape[0] and sub.shape[1]:
                raw += filt(pack(sub), bpp)

    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xFFFFFFFF)
    data = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, ctype, 0, 0, 1 if interlace else 0))
    if palette is not None:
        data += chunk(b"PLTE", palette.astype(np.uint8).tobytes())
    if trns is not None:
        data += chunk(b"tRNS", trns)
    half = len(raw) // 2 if len(raw) > 40 else len(raw)
    z = zlib.compress(raw, 6)
    data += chunk(b"IDAT", z[:len(z) // 2]) + chunk(b"IDAT", z[len(z) // 2:]) + chunk(b"IEND", b"")
    open(path, "wb").write(data)


def test_png_decoder_all_colour_types_depths_interlace_against_stb_image(host, tmp_path):
    """Every PNG colour type x bit depth x {plain, Adam7} x {no tRNS, tRNS}, random scanline filters, odd sizes —
    against the reference's stb_image (needs oracle/_ref: skipped otherwise)."""
    from oracle import ref
    import struct
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(77)
    combos = [(0, d) for d in (1, 2, 4, 8, 16)] + [(2, 8), (2, 16)] + [(3, d) for d in (1, 2, 4, 8)] + [(4, 8), (4, 16), (6, 8), (6, 16)]
    count = 0
    for ctype, depth in combos:
        n = {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}[ctype]
        for interlace in (False, True):
            for with_trns in (False, True):
                if with_trns and ctype in (4, 6):
                    continue
                w, h = int(rng.integers(1, 40)), int(rng.integers(1, 40))
                palette = None
                hi = 1 << depth
                if ctype == 3:
                    palette = rng.integers(0, 256, (hi, 3))
                samples = rng.integers(0, hi, (h, w, n)).astype(np.uint16)
                if rng.random() < 0.5 and w > 3:
                    samples[:, : w // 2] = samples[0, 0]                  # flat region: makes the tRNS colour occur
                trns = None
                if with_trns:
                    if ctype == 3:
                        trns = bytes(rng.integers(0, 256, int(rng.integers(1, hi + 1))).astype(np.uint8))
                    else:
                        trns = b"".join(struct.pack(">H", int(v)) for v in samples[0, 0])
                f = str(tmp_path / ("c%d_d%d_i%d_t%d.png" % (ctype, depth, interlace, with_trns)))
                _write_png(f, w, h, ctype, depth, samples, interlace, palette, trns, rng)
                a, b = host.decode_image(f), ref.decode_image(f)
                assert a.shape == b.shape and np.array_equal(a, b), (ctype, depth, interlace, with_trns, w, h)
                count += 1
    assert count >= 48


def _write_hdr(path, rgbe, rle, tag=b"#?RADIANCE"):
    """rgbe: uint8 [h, w, 4].  rle = new-style per-channel run-length scanlines, else flat."""
    h, w, _ = rgbe.shape
    out = tag + b"\nFORMAT=32-bit_rle_rgbe\n\n" + ("-Y %d +X %d\n" % (h, w)).encode()
    for y in range(h):
        if not rle:
            out += rgbe[y].tobytes()
            continue
        out += bytes([2, 2, w >> 8, w & 255])
        for c in range(4):
            row = rgbe[y, :, c]
            x = 0
            while x < w:
                run = 1
                while x + run < w and run < 127 and row[x + run] == row[x]:
                    run += 1
                if run >= 3:
                    out += bytes([128 + run, int(row[x])])
                    x += run
                else:
                    lit = 1
                    while x + lit < w and lit < 128 and not (x + lit + 2 < w and row[x + lit] == row[x + lit + 1] == row[x + lit + 2]):
                        lit += 1
                    out += bytes([lit]) + row[x:x + lit].tobytes()
                    x += lit
    open(path, "wb").write(out)


def test_hdr_decoder_flat_and_rle_against_stb_image(host, tmp_path):
    """Radiance .hdr: run-length and flat scanlines, widths below 8 (always flat), zero exponents, both magic
    lines — float values equal to stbi_loadf's bit for bit (needs oracle/_ref: skipped otherwise)."""
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(31)
    for i, (w, h, rle) in enumerate(((40, 7, True), (40, 7, False), (5, 9, False), (64, 3, True), (300, 4, True), (8, 8, True), (33, 2, False))):
        rgbe = rng.integers(0, 256, (h, w, 4)).astype(np.uint8)
        rgbe[:, : w // 3] = rgbe[0, 0]                      # runs
        rgbe[0, -1, 3] = 0                                  # a zero exponent
        rgbe[:, :, 3] = np.clip(rgbe[:, :, 3].astype(int) % 40 + 110, 0, 255)
        rgbe[0, -1, 3] = 0
        f = str(tmp_path / ("t%d.hdr" % i))
        _write_hdr(f, rgbe, rle, b"#?RGBE" if i == 3 else b"#?RADIANCE")
        a, b = host.decode_hdr(f), ref.decode_hdr(f)
        assert a.shape == b.shape and a.tobytes() == b.tobytes(), (w, h, rle)


def test_damaged_scene_files_are_errors_with_a_reason(host, tmp_path):
    """A scene.json that is not JSON, is cut off, or is not an object; a mesh file that is missing or cut off:
    loadScene raises with the reason (the reference prints and exit(0)s, GEMLoader.h:349-354, or reads garbage)."""
    import shutil
    rs = ref_scene("cornell-box")
    d = tmp_path / "s"
    shutil.copytree(rs.dir, d, symlinks=False)
    good = open(d / "scene.json").read()
    for text, why in (("{ this is not json", "JSON"), (good[: len(good) // 2], "JSON"), (good + " x", "trailing"), ("[1, 2]", "object")):
        open(d / "scene.json", "w").write(text)
        with pytest.raises(RuntimeError) as e:
            host.load_scene(str(d))
        assert why in str(e.value), (text[:20], str(e.value))
    open(d / "scene.json", "w").write(good)
    assert host.load_scene(str(d)).n_tris == 36
    gem = sorted(p for p in os.listdir(d) if p.endswith(".gem"))[0]
    data = open(d / gem, "rb").read()
    open(d / gem, "wb").write(data[:100])
    with pytest.raises(RuntimeError) as e:
        host.load_scene(str(d))
    assert "truncated" in str(e.value)
    # a vertex index past the vertex array (the file ends with the index list): an error naming the index,
    # where the reference reads out of bounds
    open(d / gem, "wb").write(data[:-4] + (9999).to_bytes(4, "little"))
    with pytest.raises(RuntimeError) as e:
        host.load_scene(str(d))
    assert "out of range" in str(e.value) and "9999" in str(e.value)
    os.remove(d / gem)
    with pytest.raises(RuntimeError):
        host.load_scene(str(d))


def test_truncated_rle_hdr_is_rejected_without_reading_past_the_end(host, tmp_path):
    """A run-length .hdr cut 1-3 bytes into a scanline header must fail cleanly (it used to read file[pos + 3])."""
    rng = np.random.default_rng(5)
    rgbe = rng.integers(100, 140, (6, 40, 4)).astype(np.uint8)
    f = str(tmp_path / "full.hdr")
    _write_hdr(f, rgbe, True, b"#?RADIANCE")
    assert host.decode_hdr(f).shape == (6, 40, 3)
    data = open(f, "rb").read()
    # find the start of the last scanline header (2, 2, hi, lo) and cut right after 1..3 of its bytes
    last = data.rfind(bytes([2, 2, 0, 40]))
    assert last > 0
    for keep in (1, 2, 3):
        g = str(tmp_path / ("cut%d.hdr" % keep))
        open(g, "wb").write(data[: last + keep])
        with pytest.raises((RuntimeError, ValueError)):
            host.decode_hdr(g)


def test_image_writers_round_trip(host, tmp_path):
    """Film::save's .hdr and savePNG's .png as written by the stand-alone C++ program: the PNG decodes to the
    same bytes (own decoder, and stb_image when oracle/_ref is there); the RGBE file decodes to the values
    within the format's 1/256 mantissa step, identically by both decoders."""
    from oracle import ref
    rng = np.random.default_rng(12)
    for c in (1, 3, 4):
        img = rng.integers(0, 256, (23, 37, c), dtype=np.uint8)
        f = str(tmp_path / ("w%d.png" % c))
        host.write_png(f, img)
        assert np.array_equal(host.decode_image(f), img)
        if ref.available():
            assert np.array_equal(ref.decode_image(f), img)
    rgb = (rng.random((19, 41, 3)) ** 4 * 50).astype(np.float32)
    rgb[0, 0] = 0
    rgb[1, 1] = [1e-3, 2.0, 300.0]
    f = str(tmp_path / "w.hdr")
    host.write_hdr(f, rgb)
    back = host.decode_hdr(f)
    assert np.all(np.abs(back - rgb) <= rgb.max(axis=-1, keepdims=True) / 128 + 1e-30)
    assert back[0, 0].tolist() == [0, 0, 0]
    if ref.available():
        assert back.tobytes() == ref.decode_hdr(f).tobytes()
