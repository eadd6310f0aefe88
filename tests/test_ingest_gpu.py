"""Device-side FASTA ingest (frisk_ingest.cu) against the host packer, which tests/test_host.py holds
to the reference's iterFasta/countN rules (F:139-164, F:106-118): same records, same layout, and
bit-identical planes; then the whole FASTA-text-in, rows-out path against the golden fixtures."""
import ctypes as C

import numpy as np
import pytest

from frisk_b200 import _lib, engine, synth
from tests.helpers import Golden

pytestmark = pytest.mark.gpu


def fasta_text(scaffolds, width=60, eol=b"\n", trailing_newline=True):
    out = []
    for name, seq in scaffolds:
        out.append(b">" + name.encode() + b" some description" + eol)
        b = seq.tobytes() if isinstance(seq, np.ndarray) else seq
        if width <= 0:
            out.append(b + eol)
        else:
            out.extend(b[i:i + width] + eol for i in range(0, len(b), width))
    text = b"".join(out)
    if not trailing_newline and text.endswith(eol):
        text = text[:-len(eol)]
    return text


def open_stats():
    out = np.zeros(2, dtype=np.uint64)
    _lib.check(_lib.lib().frisk_b200_fasta_open_stats(out.ctypes.data), "fasta_open_stats")
    return int(out[0]), int(out[1])


class ingest_options:
    """frisk_b200_fasta_open modes: default (chunks of >= 2 MiB), chunks of one or a few 4096-byte tiles (so that
    test-sized texts cross chunk boundaries everywhere), and the one-piece exact open."""
    MODES = {"default": {}, "chunk1": {"ingest_chunk_tiles": 1}, "chunk3": {"ingest_chunk_tiles": 3},
             "exact": {"ingest_exact_open": 1}}

    def __init__(self, mode):
        self.opts = self.MODES[mode]

    def __enter__(self):
        for k, v in self.opts.items():
            _lib.check(_lib.lib().frisk_b200_set_option(k.encode(), v), "set_option")

    def __exit__(self, *exc):
        for k in self.opts:
            _lib.lib().frisk_b200_set_option(k.encode(), 0)


def assert_same_genome(text, modes=("default", "chunk1", "chunk3", "exact")):
    host = engine.PackedGenome.from_fasta_bytes(text)
    for mode in modes:
        with ingest_options(mode):
            dev = engine.DeviceGenome.from_fasta_bytes(text).to_host()
        _assert_same(dev, host)
    return host


def _assert_same(dev, host):
    assert dev.names == host.names
    assert np.array_equal(dev.scaf_len, host.scaf_len)
    assert np.array_equal(dev.scaf_off, host.scaf_off)
    assert dev.padded_len == host.padded_len
    assert (dev.total_len, dev.nn_total, dev.n_lower) == (host.total_len, host.nn_total, host.n_lower)
    assert np.array_equal(dev.codes, host.codes)
    assert np.array_equal(dev.inv, host.inv)
    assert (dev.low is None) == (host.low is None)
    if host.low is not None:
        assert np.array_equal(dev.low, host.low)


def test_reference_scanning_rules():
    text = (b"leading junk before any header\n"
            b">>seq1 description words\nACGT\n  \nacgtNN\r\n"
            b">seq2\tother\n\nAC\nGT\n"
            b">empty_record\n"
            b"  >indented_header x\n ACGT \t\n"
            b">last\nTT>TT")
    g = assert_same_genome(text)
    assert g.names == ["seq1", "seq2", "empty_record", "indented_header", "last"]
    assert list(g.scaf_len) == [10, 4, 0, 4, 5]


@pytest.mark.parametrize("width,eol,trail", [(60, b"\n", True), (80, b"\r\n", True), (1, b"\n", False), (0, b"\n", True),
                                             (4095, b"\n", True), (4096, b"\n", False), (7, b"\r\n", False)])
def test_line_layouts(width, eol, trail):
    sc = synth.make("edge") + synth.make("C2", 0.003, seed=5)
    assert_same_genome(fasta_text(sc, width, eol, trail))


def test_headers_around_tile_boundaries():
    rng = np.random.default_rng(3)
    letters = np.frombuffer(b"ACGTacgtNRY", dtype=np.uint8)
    for shift in range(4090, 4102):
        first = letters[rng.integers(0, len(letters), shift)].tobytes()
        text = b">a\n" + first + b"\n>b\n" + b"ACGT" * 3000 + b"\n>c\n\n>d\nA\n"
        assert_same_genome(text)
    # a header line longer than two tiles (and, with test-sized chunks, longer than a chunk), one that ends exactly at a tile
    # boundary, and a '>' inside a sequence line right after such a header
    for hdr_len in (4093, 4094, 9000, 20000):
        text = b">a\nACGT\n>" + b"h" * hdr_len + b" d\n" + b"AC>GT" * 500 + b"\n>c\nTTTT\n"
        g = assert_same_genome(text)
        assert list(g.scaf_len) == [4, 2500, 4]
    # a tile made of headers only, and one made of blank lines only
    text = b"".join(b">r%05d\n" % i for i in range(2000)) + b"\n" * 9000 + b">z\nACGT\n" + b" " * 5000 + b"\nGG\n"
    g = assert_same_genome(text)
    assert len(g.names) == 2001 and int(g.scaf_len[-1]) == 6


def test_fuzzed_texts():
    rng = np.random.default_rng(11)
    alphabet = np.frombuffer(b"ACGTACGTACGTacgtNnRYKM-*>", dtype=np.uint8)
    for trial in range(12):
        parts = [b"junk line\n"] if trial % 3 == 0 else []
        for r in range(int(rng.integers(1, 60))):
            parts.append(b" " * int(rng.integers(0, 3)) + b">" * int(rng.integers(1, 3)) + b"rec%d_%d" % (trial, r) +
                         (b"\tdesc" if r % 2 else b"") + (b"\r\n" if r % 5 == 0 else b"\n"))
            n = int(rng.integers(0, 9000)) if r % 7 else 0
            seq = alphabet[rng.integers(0, len(alphabet), n)].tobytes()
            width = int(rng.choice([1, 13, 60, 61, 500, 10000]))
            for i in range(0, n, width):
                line = seq[i:i + width]
                if line.lstrip(b" \t").startswith(b">"):
                    line = b"A" + line[1:]                 # a sequence line must not look like a header
                parts.append(line + (b"\r\n" if trial % 2 else b"\n"))
                if rng.random() < 0.05:
                    parts.append(b"  \t \n")
        text = b"".join(parts)
        if trial % 4 == 1:
            text = text.rstrip(b"\r\n")
        assert_same_genome(text)


@pytest.mark.parametrize("text", [b">a\nAC GT\n", b">a\nACGT\n>b\nAC\tGT\nAA\n", b">a\n" + b"ACGT" * 2000 + b"\rA\n",
                                  b">a\n" + b"A" * 4090 + b"  \t  C\n", b">a\nAC" + b" " * 6000 + b"GT\n"])
def test_whitespace_inside_a_sequence_line_is_refused(text):
    """The reference strips a line only at its ends (F:149): whitespace INSIDE a sequence line would be a character of
    its sequence (counted in totalLen, nnTotal and every coordinate after it).  Host scanner and device ingest both
    refuse such text instead of silently deviating; whitespace at the ends of lines (CRLF, indentation) is fine."""
    from frisk_b200 import engine
    with pytest.raises(_lib.FriskError) as ei:
        engine.DeviceGenome.from_fasta_bytes(text)
    assert ei.value.code == _lib.E_FORMAT
    with pytest.raises(_lib.FriskError) as ei:
        engine.PackedGenome.from_fasta_bytes(text)
    assert ei.value.code == _lib.E_FORMAT
    assert_same_genome(b">a\n  ACGT \t \r\n\tGG\r\n" + b" " * 5000 + b"\n>b x y\nAC\n")


def test_empty_and_malformed_inputs():
    g = engine.DeviceGenome.from_fasta_bytes(b"")
    assert g.host.names == [] and g.host.padded_len == 128 and g.host.total_len == 0
    assert np.all(g.to_host().inv == 0xFFFFFFFF)
    g = engine.DeviceGenome.from_fasta_bytes(b"no header at all\nACGT\n")
    assert g.host.names == [] and g.host.total_len == 0
    with pytest.raises(_lib.FriskError) as ei:
        engine.DeviceGenome.from_fasta_bytes(b">ok\nACGT\n>\nACGT\n")
    assert ei.value.code == _lib.E_FORMAT                     # the reference raises IndexError at F:156


@pytest.mark.parametrize("case", ["edge_default", "edge_scaffoldsAll", "edge_maskHost", "c1_small", "c2_small",
                                  "c2_small_query_vs_c1_host",
                                  # other word sizes / window shapes: kmin > 1 (device tail, kmin template), kmax 9 and
                                  # kmax <= 5 (host-driven tail), windows shorter than a step's complement, scaffoldsAll rescue
                                  "edge_k4_8", "edge_k1_9", "edge_k2_5_w1000_i250", "edge_k1_3_w3000_i1000",
                                  "short_w1000_i800", "short_w1000_i800_all"])
def test_fasta_text_to_rows_matches_reference_golden(case):
    from tests.test_gpu_parity import _check_against_golden
    gold = Golden(case)
    host = gold.host()
    qtext, htext = fasta_text(gold.scaffolds()), fasta_text(host) if host is not None else None
    res = engine.run_fasta(qtext, htext, **gold.kwargs())
    _check_against_golden(res, gold, case + "[run_fasta]")
    # the streamed call with test-sized chunks (planes laid out, packed and counted chunk by chunk), the exact open behind
    # it, and a row buffer that is too small (second stage through frisk_b200_run_resident on the handles' planes)
    for mode in ("chunk1", "chunk3", "exact"):
        with ingest_options(mode):
            alt = engine.run_fasta(qtext, htext, **gold.kwargs())
        assert np.array_equal(alt.rows, res.rows, equal_nan=True) and np.array_equal(alt.tables, res.tables), mode
    # page-locked text: every chunk's copy is queued before anything else is set up (a different order of the same work)
    pq = engine._alloc(len(qtext), np.uint8, True)
    pq[:] = np.frombuffer(qtext, dtype=np.uint8)
    ph = None
    if htext is not None:
        ph = engine._alloc(len(htext), np.uint8, True)
        ph[:] = np.frombuffer(htext, dtype=np.uint8)
    for mode in ("default", "chunk1"):
        with ingest_options(mode):
            alt = engine.run_fasta(pq, ph, **gold.kwargs())
        assert np.array_equal(alt.rows, res.rows, equal_nan=True) and np.array_equal(alt.tables, res.tables), "pinned " + mode
        assert alt.names == res.names
    pageable = engine.HostOutputs(res.n_candidates + 3, gold.kwargs().get("kmax", 8), pinned=False)   # rows come back by copy
    alt = engine.run_fasta(qtext, htext, out=pageable, **gold.kwargs())
    assert np.array_equal(alt.rows, res.rows, equal_nan=True) and np.array_equal(alt.tables, res.tables)
    small = engine.HostOutputs(1, gold.kwargs().get("kmax", 8))
    alt = engine.run_fasta(qtext, htext, out=small, **gold.kwargs())
    assert np.array_equal(alt.rows, res.rows, equal_nan=True) and np.array_equal(alt.tables, res.tables)
    # same kernels as the packed-plane entry point -> identical bits
    q = engine.PackedGenome.from_scaffolds(gold.scaffolds())
    h = engine.PackedGenome.from_scaffolds(host) if host is not None else None
    ref = engine.run(q, h, **gold.kwargs())
    assert np.array_equal(res.rows, ref.rows, equal_nan=True) and np.array_equal(res.tables, ref.tables)


@pytest.mark.parametrize("kmax,w,step", [(7, 5000, 2500), (7, 1200, 600), (6, 5000, 2500), (8, 8000, 4000), (8, 9000, 4500)])
def test_fasta_text_path_equals_the_packed_plane_path(kmax, w, step):
    """Whichever tail frisk_b200_run_fasta takes -- device-built window list with the nibble kernel (kmax 7 and 8, windows up to
    8,186 bases), host-built list otherwise -- its rows are bit for bit those of the packed-plane entry point."""
    sc = synth.make("edge") + synth.make("C2", 0.004, seed=3)
    text = fasta_text(sc)
    ref = engine.run(engine.PackedGenome.from_scaffolds(sc), kmax=kmax, w=w, step=step, scaffolds_all=True)
    got = engine.run_fasta(text, kmax=kmax, w=w, step=step, scaffolds_all=True)
    assert np.array_equal(got.rows, ref.rows, equal_nan=True) and np.array_equal(got.tables, ref.tables)
    assert got.names == ref.names and np.array_equal(got.coords, ref.coords) and np.array_equal(got.status, ref.status)


def test_large_single_line_record_and_many_small_records():
    big = synth.make("C1", 0.2)                                # one 1 Mbp scaffold on a single line
    assert_same_genome(fasta_text(big, width=0))
    small = synth.make("C5", 0.00002)                          # ~20 scaffolds around 9 kbp ...
    rng = np.random.default_rng(2)
    tiny = [("t%d" % i, np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, int(rng.integers(0, 40)))])
            for i in range(5000)]                              # ... and 5,000 records shorter than a line
    assert_same_genome(fasta_text(small + tiny, width=60))


def test_chunked_open_carries_state_across_chunks_and_falls_back_when_it_cannot_decide():
    """The chunked open (text uploaded in pieces, each tokenised while the next is on the bus) carries the record count,
    the header/sequence state of the open line and the bases of the open record from chunk to chunk; what it cannot
    decide without a later chunk, or more records than its guess, is redone by the exact open.  Same genome always."""
    rng = np.random.default_rng(5)
    sc = synth.make("C2", 0.02, seed=9)                         # ~0.8 MB of text: 8 chunks of ~25 tiles with chunk1
    text = fasta_text(sc, 60)
    done0, redo0 = open_stats()
    assert_same_genome(text, modes=("chunk1",))
    done1, redo1 = open_stats()
    assert (done1 - done0, redo1 - redo0) == (1, 0), "an ordinary FASTA file completes on the chunked path"
    # one record spanning every chunk, on a single line and on many
    one = [("solo", np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, 300_000)])]
    for width in (0, 61):
        assert_same_genome(fasta_text(one, width), modes=("chunk1", "chunk3"))
    assert open_stats()[1] == redo1
    # a header line and a blank run lying across chunk boundaries (chunks of 4 tiles = 16,384 bytes each)
    for shift in (16372, 16382, 16383, 16384, 16385):
        body = b">a\n" + b"A" * (shift - 4) + b"\n"
        for sep in (b">b description\n", b"   \t  \n", b"  >c\n"):
            assert_same_genome(body + sep + b"ACGT" * 14000 + b"\n>z\nGG\n", modes=("chunk1",))
    assert open_stats()[1] > redo1, "blank bytes before a chunk boundary need the next chunk: redone exactly"
    # the one-call path queues tables, IVOM and the window kernel from inside the streamed open; when that open is redone,
    # what was queued on the provisional data is discarded and the host-driven tail produces the same rows
    amb = b">a\n" + b"A" * (16384 - 8) + b"\n" + b"   \t  \n" + b"ACGT" * 14000 + b"\n>z\n" + b"GATTACA" * 700 + b"\n"
    with ingest_options("exact"):
        want = engine.run_fasta(amb, w=1000, step=500)
    d, r = open_stats()
    with ingest_options("chunk1"):
        got = engine.run_fasta(amb, w=1000, step=500)
    assert open_stats()[1] == r + 1, "this text cannot be decided chunk by chunk"
    assert np.array_equal(got.rows, want.rows, equal_nan=True) and np.array_equal(got.tables, want.tables) and len(got.rows) > 50
    assert got.names == want.names and np.array_equal(got.coords, want.coords)
    # more records than the guess (n/64 + 4096)
    done2, redo2 = open_stats()
    many = b"".join(b">r%d\nA\n" % i for i in range(60_000))
    g = assert_same_genome(many, modes=("default",))
    assert len(g.names) == 60_000 and open_stats()[1] == redo2 + 1
    with ingest_options("exact"):
        d, r = open_stats()
        engine.DeviceGenome.from_fasta_bytes(text)
        assert open_stats() == (d, r), "the exact open is neither chunked nor a retry"


@pytest.mark.parametrize("w,step,all_", [(5000, 2500, False), (5000, 2500, True), (1000, 800, False), (1000, 800, True),
                                        (700, 1000, True), (64, 1, False), (3000, 3000, True)])
def test_device_window_list_is_the_host_window_list(w, step, all_):
    """frisk_b200_run_fasta builds its window list on the device (frisk_windows.cu) so that nothing between the last byte of the
    text and the window kernel waits for the host: same windows, same order as frisk_b200_windows (crawlGenome, F:194-251) --
    size rule, grid, jump-back tail, --scaffoldsAll rescue, scaffolds shorter than a window or a step."""
    import torch
    rng = np.random.default_rng(w + step)
    lens = np.concatenate([rng.integers(0, 4 * w, 300), rng.integers(0, 40, 50), np.array([w, w - 1, w + 1, step, step - 1, 2 * w,
                           int(1.75 * w - step), int(1.75 * w - step) + 1, 0, 1, 250_000])]).astype(np.uint64)
    rng.shuffle(lens)
    off = np.zeros(len(lens), np.uint64)
    padded = C.c_uint64(0)
    L = _lib.lib()
    _lib.check(L.frisk_b200_pack_layout(engine._ptr(lens), len(lens), engine._ptr(off), C.byref(padded)), "pack_layout")
    n = C.c_uint64(0)
    _lib.check(L.frisk_b200_windows(engine._ptr(lens), engine._ptr(off), len(lens), w, step, int(all_), 0, None, None, None, None,
                                    None, C.byref(n)), "windows")
    cap = int(n.value)
    h_off, h_len = np.zeros(cap, np.uint64), np.zeros(cap, np.uint32)
    _lib.check(L.frisk_b200_windows(engine._ptr(lens), engine._ptr(off), len(lens), w, step, int(all_), cap, engine._ptr(h_off),
                                    engine._ptr(h_len), None, None, None, C.byref(n)), "windows")
    dev = torch.device("cuda:0")
    d_lens, d_off = torch.from_numpy(lens.view(np.int64)).to(dev), torch.from_numpy(off.view(np.int64)).to(dev)
    for dcap in (cap + 7, max(cap - 5, 0)):
        d_wo = torch.full((dcap + 1,), -1, dtype=torch.int64, device=dev)
        d_wl = torch.full((dcap + 1,), -1, dtype=torch.int32, device=dev)
        d_n = torch.zeros(2, dtype=torch.int64, device=dev)
        _lib.check(L.frisk_b200_windows_device(engine._ptr(d_lens), engine._ptr(d_off), len(lens), w, step, int(all_), dcap,
                                               engine._ptr(d_wo), engine._ptr(d_wl), C.c_void_p(d_n.data_ptr()),
                                               C.c_void_p(d_n.data_ptr() + 8), engine._stream_ptr(dev)), "windows_device")
        torch.cuda.synchronize()
        assert int(d_n[0]) == cap and int(d_n[1]) == int(lens.sum())
        k = min(cap, dcap)
        assert np.array_equal(d_wo.cpu().numpy()[:k].view(np.uint64), h_off[:k])
        assert np.array_equal(d_wl.cpu().numpy()[:k].view(np.uint32), h_len[:k])
        assert int(d_wo[dcap]) == -1 and int(d_wl[dcap]) == -1, "nothing is written beyond the capacity"


def test_multi_gigabyte_text_beyond_2_pow_32_bytes():
    """C4-scale input: 4.5 GB of FASTA text (55 k records, 4.4 Gbp) through the device ingest -- byte offsets,
    record positions and packed positions beyond 2^32 -- against the host packer, plane for plane."""
    import torch
    block = np.frombuffer(synth.fasta_bytes(synth.make("C2", 1.0, seed=77)), dtype=np.uint8)
    reps = (4_500_000_000 + len(block) - 1) // len(block)
    text = np.tile(block, reps)
    assert text.shape[0] > 2 ** 32
    dg = engine.DeviceGenome.from_fasta_bytes(text)
    host = engine.PackedGenome.from_fasta_bytes(text)
    assert dg.host.names == host.names and len(host.names) == 500 * reps
    assert np.array_equal(dg.host.scaf_len, host.scaf_len) and np.array_equal(dg.host.scaf_off, host.scaf_off)
    assert dg.host.padded_len == host.padded_len > 2 ** 32
    assert (dg.host.total_len, dg.host.nn_total, dg.host.n_lower) == (host.total_len, host.nn_total, host.n_lower)
    dev = dg.device
    assert torch.equal(dg.codes, torch.from_numpy(host.codes.view(np.int32)).to(dev))
    assert torch.equal(dg.inv, torch.from_numpy(host.inv.view(np.int32)).to(dev))
    assert (dg.low is None) == (host.low is None)
    del dg, host, text
    torch.cuda.empty_cache()
