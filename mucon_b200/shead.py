"""The sequence-generation ("s") head at test time on the GPU -- host side.

Drop-in surface: `SHead` holds the s-head's parameters under the reference's attribute names (`fs_encoder_lstm`,
`fs_encoder_hidden_out`, `fs_encoder_cn_out`, `fs_decoder_attention_W1` / `_l2` / `_l3` / `_V`, `fs_decoder_embedding`,
`fs_decoder_attn_combine`, `fs_decoder_lstm`, `fs_decoder_transcript`, `fs_decoder_length`; reference
src/mucon/models.py:193-273), so the `fs_*` entries of a reference state_dict load unchanged, and runs
`sequence_generation_forward` (models.py:585-728) for a packed batch of variable-length videos in four launches:
two input-projection GEMMs (mucon_conv1d), the BiLSTM recurrence (mucon_lstm_encoder), the attention-projection GEMM
and the attention decoder with all its steps (mucon_seq_decoder).  Inference only (eval mode: dropout off); the greedy
loop's argmax / EOS test / next input stay on the device (the reference does a `.item()` per step, models.py:721).
"""
import ctypes as C

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .temporal import _stream, conv1d_rows


class SHead(nn.Module):
    def __init__(self, num_classes=48, hidden_size=128, max_decoding_steps=31):
        super().__init__()
        if hidden_size != 128:
            raise NotImplementedError("the s-head kernels are built for hidden_size = 128")
        H = hidden_size
        self.num_classes, self.hidden_size, self.max_decoding_steps = num_classes, H, max_decoding_steps
        self.EOS_token_id = num_classes                                        # models.py:151
        self.fs_encoder_lstm = nn.LSTM(input_size=H, hidden_size=H, batch_first=True, bidirectional=True)
        self.fs_encoder_hidden_out = nn.Linear(2 * H, H)
        self.fs_encoder_cn_out = nn.Linear(2 * H, H)
        self.fs_decoder_attention_W1 = nn.Parameter(torch.randn(2 * H, H) * 0.1)
        self.fs_decoder_attention_l2 = nn.Linear(H, H)
        self.fs_decoder_attention_l3 = nn.Linear(2 * H, H)                     # defined by the reference, never used
        self.fs_decoder_attention_V = nn.Parameter(torch.randn(H) * 0.1)
        self.fs_decoder_embedding = nn.Embedding(num_classes + 2, H)
        self.fs_decoder_attn_combine = nn.Linear(3 * H, H)
        self.fs_decoder_lstm = nn.LSTM(input_size=H, hidden_size=H)
        self.fs_decoder_transcript = nn.Sequential(nn.Linear(H, H), nn.ReLU(), nn.Linear(H, num_classes + 1))
        self.fs_decoder_length = nn.Sequential(nn.Linear(H + num_classes + 1, H // 2), nn.ReLU(), nn.Linear(H // 2, 1))
        self._cache = None

    def _weights(self):
        key = tuple(p._version for p in self.parameters()) + (str(self.fs_decoder_attention_V.device),)
        if self._cache is None or self._cache[0] != key:
            f = lambda t: t.detach().contiguous().float()
            e = self.fs_encoder_lstm
            keep = dict(
                wih_f=f(e.weight_ih_l0).t().contiguous()[None], wih_b=f(e.weight_ih_l0_reverse).t().contiguous()[None],
                b_f=f(e.bias_ih_l0 + e.bias_hh_l0), b_b=f(e.bias_ih_l0_reverse + e.bias_hh_l0_reverse),
                whh_f=f(e.weight_hh_l0), whh_b=f(e.weight_hh_l0_reverse),
                w1=f(self.fs_decoder_attention_W1)[None], zero_h=torch.zeros(self.hidden_size, device=e.weight_hh_l0.device),
                hid_w=f(self.fs_encoder_hidden_out.weight), hid_b=f(self.fs_encoder_hidden_out.bias),
                cn_w=f(self.fs_encoder_cn_out.weight), cn_b=f(self.fs_encoder_cn_out.bias),
                l2_w=f(self.fs_decoder_attention_l2.weight), l2_b=f(self.fs_decoder_attention_l2.bias),
                att_v=f(self.fs_decoder_attention_V), emb=f(self.fs_decoder_embedding.weight),
                comb_w=f(self.fs_decoder_attn_combine.weight), comb_b=f(self.fs_decoder_attn_combine.bias),
                wih=f(self.fs_decoder_lstm.weight_ih_l0), whh=f(self.fs_decoder_lstm.weight_hh_l0),
                bih=f(self.fs_decoder_lstm.bias_ih_l0), bhh=f(self.fs_decoder_lstm.bias_hh_l0),
                t1_w=f(self.fs_decoder_transcript[0].weight), t1_b=f(self.fs_decoder_transcript[0].bias),
                t2_w=f(self.fs_decoder_transcript[2].weight), t2_b=f(self.fs_decoder_transcript[2].bias),
                n1_w=f(self.fs_decoder_length[0].weight), n1_b=f(self.fs_decoder_length[0].bias),
                n2_w=f(self.fs_decoder_length[2].weight), n2_b=f(self.fs_decoder_length[2].bias))
            ws = _lib.SHeadWeights()
            for name, _ in _lib.SHeadWeights._fields_:
                setattr(ws, name, keep[name].data_ptr())
            self._cache = (key, keep, ws)
        return self._cache[1], self._cache[2]

    def forward_packed(self, z, row_off, row_off_host, transcripts_tf_input=None, teacher_forcing=True, max_steps=None):
        """z [sum Tz, 128] CUDA float32 (temporal_modeling_forward's output, videos concatenated), row_off [V+1] int64
        CUDA offsets (+ the same on the host).  teacher_forcing=True: transcripts_tf_input = per video [N+1] ints
        (SOS + transcript, general_dataset's transcript_tf_input) and N+1 steps are decoded; False: greedy decoding
        from SOS (= num_classes + 1) until EOS (= num_classes) or max_steps.
        -> dict(logp [V, S, C+1], lengths [V, S], tokens [V, S], n_steps [V], encoder_out [sum Tz, 256])."""
        if self.training:
            raise NotImplementedError("the s-head kernels are inference-only; call .eval()")
        if not z.is_cuda:
            raise _lib.MuconError("the s-head needs CUDA tensors (there is no CPU fallback)")
        w, ws = self._weights()
        H, dev = self.hidden_size, z.device
        z = z.detach().contiguous().float()
        V = int(row_off_host.shape[0]) - 1
        Tz = np.diff(np.asarray(row_off_host, dtype=np.int64))
        max_Tz = int(Tz.max(initial=0))
        lib = _lib.lib()
        if V == 0:
            e = lambda *shape, dt=torch.float32: torch.zeros(shape, dtype=dt, device=dev)
            return dict(logp=e(0, 1, self.num_classes + 1), lengths=e(0, 1), tokens=e(0, 1, dt=torch.int32),
                        n_steps=e(0, dt=torch.int32), encoder_out=e(0, 2 * H))
        if z.shape[0] == 0:
            z = torch.zeros((1, H), dtype=torch.float32, device=dev)   # no pooled rows at all: keep the launches valid
        def sgemm(A, B, bias):   # exact fp32 (mucon_sgemm_bias): A [M, K] . B [K, N] + bias
            out = torch.empty((A.shape[0], B.shape[1]), dtype=torch.float32, device=dev)
            _lib.check(lib.mucon_sgemm_bias(_lib.ptr(A), _lib.ptr(B), _lib.ptr(bias), _lib.ptr(out), C.c_int64(A.shape[0]),
                                            C.c_int(A.shape[1]), C.c_int(B.shape[1]), _stream(dev)), "mucon_sgemm_bias")
            return out
        xp_f = sgemm(z, w["wih_f"][0], w["b_f"])                                    # [rows, 512]
        xp_b = sgemm(z, w["wih_b"][0], w["b_b"])
        enc = torch.empty((z.shape[0], 2 * H), dtype=torch.float32, device=dev)
        hn = torch.empty((V, 2, H), dtype=torch.float32, device=dev)
        cn = torch.empty((V, 2, H), dtype=torch.float32, device=dev)
        order = torch.from_numpy(np.argsort(-Tz, kind="stable").astype(np.int32)).to(dev)   # similar lengths share a CTA
        _lib.check(lib.mucon_lstm_encoder(_lib.ptr(xp_f), _lib.ptr(xp_b), _lib.ptr(w["whh_f"]), _lib.ptr(w["whh_b"]),
                                          _lib.ptr(row_off), _lib.ptr(order), C.c_int(V), C.c_int(H), _lib.ptr(enc),
                                          _lib.ptr(hn), _lib.ptr(cn), _stream(dev)), "mucon_lstm_encoder")
        enc_ready = sgemm(enc, w["w1"][0], None)                                     # [rows, 128] (models.py:627-629)
        if teacher_forcing:
            if transcripts_tf_input is None:
                raise ValueError("teacher forcing needs transcripts_tf_input")
            tf = [np.asarray(t, dtype=np.int32).reshape(-1) for t in transcripts_tf_input]
            S = max(int(max((t.shape[0] for t in tf), default=1)), 1)
        else:
            tf = [np.array([self.num_classes + 1], dtype=np.int32)] * V           # SOS
            S = int(max_steps or self.max_decoding_steps)
        tf_off = np.concatenate([[0], np.cumsum([t.shape[0] for t in tf])]).astype(np.int32)
        meta = torch.from_numpy(np.concatenate([tf_off, np.concatenate(tf) if V else np.zeros(0, np.int32)])).to(dev)
        tf_off_d, tf_in_d = meta[:V + 1], meta[V + 1:]
        nw = self.num_classes + 1
        logp = torch.zeros((V, S, nw), dtype=torch.float32, device=dev)
        lens = torch.zeros((V, S), dtype=torch.float32, device=dev)
        toks = torch.full((V, S), -1, dtype=torch.int32, device=dev)
        nst = torch.zeros(V, dtype=torch.int32, device=dev)
        # four videos share a CTA: group by the number of steps (teacher forcing) / by length (greedy)
        key = np.array([t.shape[0] for t in tf]) * 100000 + Tz if teacher_forcing else Tz
        dorder = torch.from_numpy(np.argsort(-key, kind="stable").astype(np.int32)).to(dev)
        _lib.check(lib.mucon_seq_decoder(
            C.byref(ws), _lib.ptr(enc), _lib.ptr(enc_ready), _lib.ptr(hn), _lib.ptr(cn), _lib.ptr(row_off),
            _lib.ptr(dorder), C.c_int(V),
            C.c_int(max_Tz), _lib.ptr(tf_in_d), _lib.ptr(tf_off_d), C.c_int(int(bool(teacher_forcing))), C.c_int(S),
            C.c_int(nw), C.c_int(self.EOS_token_id), _lib.ptr(logp), _lib.ptr(lens), _lib.ptr(toks), _lib.ptr(nst),
            _stream(dev)), "mucon_seq_decoder")
        return dict(logp=logp, lengths=lens, tokens=toks, n_steps=nst, encoder_out=enc)

    def sequence_generation_forward(self, temporal_encoded, tf_transcript_target_length, transcript_tf_input,
                                    transcript_tf_target=None, teacher_forcing=True):
        """The reference signature for one video (models.py:585-598): temporal_encoded [1, Tz, 128] ->
        (list of [1, C+1] log-probabilities, list of scalar length logits)."""
        z = temporal_encoded[0]
        Tz = int(z.shape[0])
        off_h = np.array([0, Tz], dtype=np.int64)
        off = torch.from_numpy(off_h).to(z.device)
        tf = [transcript_tf_input.detach().cpu().numpy()[:tf_transcript_target_length]] if teacher_forcing else None
        out = self.forward_packed(z, off, off_h, tf, teacher_forcing=teacher_forcing)
        n = int(out["n_steps"][0].item())
        return [out["logp"][0, s:s + 1] for s in range(n)], [out["lengths"][0, s] for s in range(n)]
