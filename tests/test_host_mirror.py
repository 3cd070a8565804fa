"""Host-side mirror of the reference's Python surface (no GPU): names, layouts, error behaviour."""
import json
import os

import numpy as np
import pytest
import torch

from laughter_detection_icsi_b200 import config, datasets, laugh_segmenter, load_data, models, segment_laughter, synth, textgrid
from laughter_detection_icsi_b200._native import LdError
from laughter_detection_icsi_b200.utils import audio_utils, torch_utils
from laughter_detection_icsi_b200.utils.utils import B200Fbank, get_feat_extractor
from oracle import resnet_oracle


def test_config_matches_reference():
    base = config.MODEL_MAP["resnet_base"]
    assert base["model"] is models.ResNetBigger
    assert (base["batch_size"], base["log_frequency"], base["linear_layer_size"], base["filter_sizes"]) == (32, 900, 48, [64, 32, 16, 16])
    assert config.FEAT == {"num_samples": 100, "num_filters": 44}
    assert config.MODEL_MAP["resnet_with_augmentation"]["linear_layer_size"] == 128


def test_state_dict_layout_is_the_references(golden_dir, capsys):
    m = models.ResNetBigger(dropout_rate=0.0, linear_layer_size=48, filter_sizes=[64, 32, 16, 16])
    assert "training with dropout=0.0" in capsys.readouterr().out
    with open(os.path.join(golden_dir, "resnet_state_dict_keys.json")) as f:
        ref = json.load(f)
    sd = m.state_dict()
    assert list(sd.keys()) == list(ref.keys())
    assert all(list(sd[k].shape) == ref[k] for k in ref)
    assert (m.global_step, m.epoch, m.best_val_loss) == (0, 0, np.inf)
    m.load_state_dict(resnet_oracle.random_state_dict(seed=1))  # a reference-layout checkpoint loads unchanged
    assert synth.param_shapes() == resnet_oracle.param_shapes()


def test_forward_fails_loudly_off_gpu():
    m = models.ResNetBigger(dropout_rate=0.0, linear_layer_size=48, filter_sizes=[64, 32, 16, 16]).eval()
    with pytest.raises(LdError):
        m(torch.zeros(2, 1, 100, 44))
    m.train()
    with pytest.raises(LdError):
        m(torch.zeros(2, 1, 100, 44))


def test_checkpoint_round_trip(tmp_path):
    m = models.ResNetBigger(dropout_rate=0.0, linear_layer_size=48, filter_sizes=[64, 32, 16, 16])
    torch_utils.init_weights(m)
    assert abs(float(m.bn1.weight.std()) - 0.01) < 5e-3  # every parameter ~ N(0, 0.01), BatchNorm affine included
    opt = torch.optim.Adam(m.parameters())
    state = torch_utils.make_state_dict(m, opt, epoch=2, global_step=10, best_val_loss=0.5)
    assert set(state) == {"epoch", "global_step", "best_val_loss", "state_dict", "optim_dict"}
    torch_utils.save_checkpoint(state, True, str(tmp_path / "ck"))
    assert (tmp_path / "ck" / "last.pth.tar").exists() and (tmp_path / "ck" / "best.pth.tar").exists()
    m2 = models.ResNetBigger(dropout_rate=0.0, linear_layer_size=48, filter_sizes=[64, 32, 16, 16])
    torch_utils.load_checkpoint(str(tmp_path / "ck" / "best.pth.tar"), m2)
    assert (m2.epoch, m2.global_step, m2.best_val_loss) == (2, 11, 0.5)
    assert all(torch.equal(a, b) for a, b in zip(m.state_dict().values(), m2.state_dict().values()))


def test_inference_dataset_windows_and_padding():
    feats = np.arange(130 * 44, dtype=np.float32).reshape(130, 44)
    ds = datasets.InferenceDataset(feats)
    assert len(ds) == 130
    assert ds[0].shape == (100, 44) and np.array_equal(ds[5], feats[5:105])
    tail = ds[100]
    assert tail.shape == (100, 44) and np.array_equal(tail[:30], feats[100:]) and np.all(tail[30:] == 0)
    assert np.array_equal(tail, resnet_oracle.window(feats, 100))


def test_lad_dataset_batch_layout():
    cuts = [datasets.FeatureCut(np.full((100, 44), i, dtype=np.float32), i % 2) for i in range(5)]
    batch = datasets.LadDataset()[cuts]
    assert batch["inputs"].shape == (5, 100, 44) and batch["inputs"].dtype == torch.float32
    assert batch["is_laugh"].dtype == torch.int32 and batch["is_laugh"].tolist() == [0, 1, 0, 1, 0]
    assert batch["input_lens"].tolist() == [100] * 5 and batch["cut"] is cuts


def test_training_dataloader_split_validation():
    with pytest.raises(ValueError, match="Unexpected value for split"):
        load_data.create_training_dataloader("/nonexistent", "validation")


def test_feat_extractor_surface():
    ex = get_feat_extractor(num_samples=100, num_filters=44)
    assert isinstance(ex, B200Fbank) and ex.frame_shift == 0.01 and ex.num_filters == 44
    with pytest.raises(AssertionError):
        ex.extract(np.zeros(16000, dtype=np.float32), 8000)
    x = np.array([-1.0, 0.5, 32767 / 32768.0, 1.0])
    assert B200Fbank.to_int16(x).tolist() == [-32768, 16384, 32767, 32767]


def test_cli_flags_and_errors(tmp_path):
    p = segment_laughter.build_parser()
    a = p.parse_args(["--input_audio_file", "x.wav", "--thresholds", "0.2,0.5", "--min_length", "0.1"])
    assert a.thresholds == "0.2,0.5" and a.min_lengths == "0.1" and a.config == "resnet_with_augmentation"
    assert a.save_to_audio_files == "True" and a.save_to_textgrid == "False"
    a = p.parse_args(["--input_audio_file", "x.wav", "--threshold", "0.7"])
    assert a.thresholds == "0.7"
    with pytest.raises(SystemExit):
        p.parse_args([])
    assert segment_laughter.strtobool("True") == 1 and segment_laughter.strtobool("false") == 0
    with pytest.raises(Exception, match="Model checkpoint not found"):
        segment_laughter.load_model(str(tmp_path / "missing"), config.MODEL_MAP["resnet_base"], "cpu")


def test_textgrid_writer_round_trip(tmp_path):
    inst = [(0.5, 1.25), (3.0, 4.0)]
    path = tmp_path / "chan0.TextGrid"
    textgrid.write_laughter_textgrid(str(path), inst)
    text = path.read_text().split("\n")
    assert text[:3] == ['File type = "ooTextFile"', 'Object class = "TextGrid"', ""]
    assert text[3:12] == ["0.0", "4.0", "<exists>", "1", '"IntervalTier"', '"laughter"', "0.0", "4.0", "4"]
    ivs = textgrid.read_intervals(str(path))
    assert [(s, e) for s, e, t in ivs if t == "laugh"] == inst
    assert [t for _, _, t in ivs] == ["", "laugh", "", "laugh"]


def test_wav_helpers(tmp_path):
    import scipy.io.wavfile
    pcm = synth.synth_channel(16000 * 2 + 5).numpy()
    path = str(tmp_path / "a.wav")
    scipy.io.wavfile.write(path, 16000, pcm)
    got, sr = audio_utils.load_wav_int16(path)
    assert sr == 16000 and np.array_equal(got, pcm)
    assert audio_utils.get_audio_length(path) == len(pcm) / 16000.0


def test_segmenter_helpers():
    assert laugh_segmenter.frame_span_to_time_span((3, 7), fps=100.0) == (0.03, 0.07)
    assert laugh_segmenter.collapse_to_start_and_end_frame([4, 5, 6]) == (4, 6)
    assert laugh_segmenter.comparison_thresholds([0.3], False) == [0.3]
    assert laugh_segmenter.format_outputs([(0.1, 0.2)], ["a.wav"]) == [{"filename": "a.wav", "start": 0.1, "end": 0.2}]
    thr, ml = synth.eval_grid()
    assert len(thr) == 29 and ml == [0.0, 0.1, 0.2] and thr[0] == 0.0 and thr[-1] == 1.0


def test_synthetic_audio_is_reproducible():
    a = synth.synth_channel(16000, meeting=2, channel=1)
    b = synth.synth_channel(16000, meeting=2, channel=1)
    c = synth.synth_channel(16000, meeting=2, channel=2)
    assert a.dtype == torch.int16 and torch.equal(a, b) and not torch.equal(a, c)


def test_feature_store_cuts_follow_truncate_and_pad(tmp_path):
    from laughter_detection_icsi_b200 import compute_features as cf
    store = cf.FeatureStore()
    feats = np.arange(3000 * 44, dtype=np.float32).reshape(3000, 44)
    store.add_features("Bmr021", "chan3", feats, "Bmr021/chan3.sph", 30.0)
    c = store.cut("Bmr021", "chan3", sub_start=14.785, sub_duration=1.0, label=0)          # 1478.5 -> frame 1479 (half up)
    assert c.load_features().shape == (100, 44) and np.array_equal(c.load_features(), feats[1479:1579])
    short = store.cut("Bmr021", "chan3", sub_start=2.0, sub_duration=0.37, label=1)         # 37 frames + 63 pad frames
    w = short.load_features()
    assert np.array_equal(w[:37], feats[200:237]) and np.all(w[37:] == np.float32(cf.LOG_EPSILON)) and short.supervisions[0].custom["is_laugh"] == 1
    tail = store.cut("Bmr021", "chan3", sub_start=29.5, sub_duration=1.0, label=0)          # runs off the end of the track
    assert np.array_equal(tail.load_features()[:50], feats[2950:]) and np.all(tail.load_features()[50:] == np.float32(cf.LOG_EPSILON))
    # data-frame rows in the reference's sample_df.csv format -> shuffled cuts -> LadDataset batches of 32
    csv_path = tmp_path / "train_df.csv"
    csv_path.write_text("start,duration,sub_start,sub_duration,audio_path,meeting_id,chan_id,label\n" +
                        "".join(f"{i}.0,1.5,{i}.25,1.0,Bmr021/chan3.sph,Bmr021,chan3,{i % 2}\n" for i in range(1, 28)) +
                        "".join(f"{i}.0,0.5,{i}.1,0.5,Bmr021/chan3.sph,Bmr021,chan3,1\n" for i in range(1, 14)))
    rows = cf.read_data_df(str(csv_path))
    cuts = cf.cuts_from_dataframe(rows, store, shuffle_seed=0)
    assert len(cuts) == 40 and sorted(c.id for c in cuts) == sorted(f"cut_{i}" for i in range(40))
    batches = list(cf.training_batches(cuts))
    assert [b["inputs"].shape[0] for b in batches] == [32, 8] and batches[0]["inputs"].shape[1:] == (100, 44)
    assert batches[0]["is_laugh"].dtype == torch.int32
    # save / load round trip of the raw store format
    store.save(str(tmp_path / "store"))
    again = cf.FeatureStore.load(str(tmp_path / "store"))
    assert np.array_equal(again.tracks["Bmr021/chan3"], feats) and again.meta["Bmr021/chan3"]["num_frames"] == 3000


def test_sphere_pcm_ingest(tmp_path):
    """NIST SPHERE with uncompressed 16-bit PCM (what sph2pipe-unpacked ICSI channels look like), both byte orders; the
    shorten-compressed original is refused with a clear message."""
    pcm = synth.synth_channel(16000 + 7).numpy()

    def write(path, coding, byte_format, data):
        head = ("NIST_1A\n   1024\nsample_count -i %d\nsample_rate -i 16000\nchannel_count -i 1\nsample_n_bytes -i 2\n"
                "sample_byte_format -s2 %s\nsample_coding -s%d %s\nend_head\n" % (len(pcm), byte_format, len(coding), coding)).encode()
        with open(path, "wb") as f:
            f.write(head + b" " * (1024 - len(head)))
            f.write(data)
    le, be, shn = str(tmp_path / "le.sph"), str(tmp_path / "be.sph"), str(tmp_path / "shn.sph")
    write(le, "pcm", "01", pcm.astype("<i2").tobytes())
    write(be, "pcm", "10", pcm.astype(">i2").tobytes())
    write(shn, "pcm,embedded-shorten-v2.00", "01", b"\x00" * 100)
    for p in (le, be):
        got, sr = audio_utils.load_wav_int16(p)
        assert sr == 16000 and got.dtype == np.int16 and np.array_equal(got, pcm)
        assert audio_utils.get_audio_length(p) == len(pcm) / 16000.0
    with pytest.raises(ValueError, match="ajkg"):   # announced as shorten, but the payload is not a shorten stream
        audio_utils.load_wav_int16(shn)
    other = str(tmp_path / "ulaw.sph")
    write(other, "ulaw", "01", b"\x00" * 100)
    with pytest.raises(ValueError, match="sph2pipe"):
        audio_utils.load_wav_int16(other)


def test_shorten_decoder_round_trips_an_independent_encoder(tmp_path):
    """ICSI's channels are shorten-compressed SPHERE files.  No shorten binary or ICSI file is available offline (parity
    unpinned), so the C decoder (ld_shorten_decode) is exercised against tests/shorten_encoder.py on every block command:
    DIFF0-3, QLPC, ZERO, BITSHIFT, BLOCKSIZE (short last block), VERBATIM, running-mean offsets, mono and stereo, versions 1-3."""
    import ctypes
    import shorten_encoder as se
    from laughter_detection_icsi_b200 import _native
    lib = _native.load_library()

    def decode(stream):
        data = np.frombuffer(stream, dtype=np.uint8)
        n, ch = ctypes.c_int64(0), ctypes.c_int32(0)
        assert lib.ld_shorten_decode(data.ctypes.data, data.size, None, 0, ctypes.byref(ch), ctypes.byref(n)) == 0, lib.ld_shorten_last_error()
        out = np.empty(n.value, dtype=np.int16)
        assert lib.ld_shorten_decode(data.ctypes.data, data.size, out.ctypes.data, out.size, ctypes.byref(ch), ctypes.byref(n)) == 0
        return out, ch.value

    rng = np.random.default_rng(7)
    speech = synth.synth_channel(16000 + 77, meeting=8).numpy().astype(np.int64)
    walk = np.clip(np.cumsum(rng.integers(-300, 301, 5000)) + 1500, -32768, 32767)      # DC offset: exercises the running mean
    cmds = [se.FN_DIFF0, se.FN_DIFF1, se.FN_DIFF2, se.FN_DIFF3, se.FN_QLPC]
    for version in (1, 2, 3):
        for nmean in (0, 4):
            plan = lambda b, c: (cmds[(b + c) % 5], 0, [21, -9] if cmds[(b + c) % 5] == se.FN_QLPC else None)
            for sig in (speech, walk):
                got, ch = decode(se.encode([sig], version=version, nmean=nmean, plan=plan, verbatim=b"RIFFhdr"))
                assert ch == 1 and np.array_equal(got, sig.astype(np.int16)), (version, nmean)
    # stereo, a silent (ZERO) block, a bit-shifted stretch (samples multiples of 4), block size 64
    left = walk[:1000].copy(); left[256:320] = 0
    right = (rng.integers(-2000, 2000, 1000) // 4) * 4
    plan = lambda b, c: ((se.FN_ZERO if (c == 0 and b == 4) else se.FN_DIFF2), (2 if c == 1 else 0), None)
    got, ch = decode(se.encode([left, right], version=2, blocksize=64, nmean=4, plan=plan))
    assert ch == 2 and np.array_equal(got.reshape(-1, 2)[:, 0], left.astype(np.int16)) and np.array_equal(got.reshape(-1, 2)[:, 1], right.astype(np.int16))
    # malformed streams fail loudly
    bad = np.frombuffer(b"ajkg\x02" + b"\x00" * 64, dtype=np.uint8)
    n = ctypes.c_int64(0)
    assert lib.ld_shorten_decode(bad.ctypes.data, bad.size, None, 0, None, ctypes.byref(n)) != 0
    # the SPHERE loader: header + embedded shorten stream, count and checksum verified
    pcm = speech.astype(np.int16)
    stream = se.encode([speech], version=2, nmean=4)
    checksum = int(pcm.view(np.uint16).astype(np.uint64).sum() & 0xFFFF)
    head = ("NIST_1A\n   1024\nsample_count -i %d\nsample_rate -i 16000\nchannel_count -i 1\nsample_n_bytes -i 2\n"
            "sample_byte_format -s2 01\nsample_coding -s26 pcm,embedded-shorten-v2.00\nsample_checksum -i %d\nend_head\n" % (len(pcm), checksum)).encode()
    path = str(tmp_path / "chan0.sph")
    with open(path, "wb") as f:
        f.write(head + b" " * (1024 - len(head)) + stream)
    got, sr = audio_utils.load_wav_int16(path)
    assert sr == 16000 and np.array_equal(got, pcm) and audio_utils.get_audio_length(path) == len(pcm) / 16000.0
    with open(path, "wb") as f:
        f.write(head.replace(b"sample_checksum -i %d" % checksum, b"sample_checksum -i %d" % ((checksum + 1) & 0xFFFF)) + b" " * 1024)
        f.seek(1024); f.write(stream)
    with pytest.raises(ValueError, match="checksum"):
        audio_utils.load_wav_int16(path)
