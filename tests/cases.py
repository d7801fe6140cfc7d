"""Seeded cases shared by the CPU and GPU suites; they regenerate the inputs of tests/golden/*.npz
(the fixtures hold the REFERENCE outputs, written by oracle/make_golden.py)."""
ACOUSTIC_CASES = {
    "ac_small": (dict(seed=11, B=3, min_chars=3, max_chars=7, max_frames=48, Lk_cap=40), False),
    "ac_ragged": (dict(seed=12, B=4, min_chars=1, max_chars=9, max_frames=64, Lk_cap=64, pron_modified_p=0.1), False),
    "ac_preddur": (dict(seed=13, B=3, min_chars=2, max_chars=6, max_frames=40, Lk_cap=32), True),
    "ac_single": (dict(seed=14, B=1, min_chars=12, max_chars=12, max_frames=120, Lk_cap=96), False),
}
VOCODER_CASES = {"voc_small": dict(seed=21, B=2, T=24), "voc_single": dict(seed=22, B=1, T=57)}
ACOUSTIC_SEED = 1234
VOCODER_SEED = 4321

# north-star tolerances (BASELINE.json): mel <= 1e-3 max-abs, wav <= 1e-4 RMS, integers bit-exact
TOL_MEL_MAXABS = 1e-3
TOL_WAV_RMS = 1e-4

# PortaSpeech sibling: tests/golden/ps_full.npz holds the reference outputs for these seeds
# (oracle/make_golden_ps.py: FULL_BATCH / FULL_WEIGHT_SEED / FULL_PH_SIZE)
PS_FULL_BATCH = dict(seed=31, B=4, min_words=3, max_words=9, max_ph_per_word=4, max_frames=64)
PS_FULL_WEIGHT_SEED = 2468
PS_FULL_PH_SIZE = 80
PS_STAGES = ("ph_encoder_out", "word_encoder_out", "dur", "attn", "decoder_inp", "z_p", "mel_out")
