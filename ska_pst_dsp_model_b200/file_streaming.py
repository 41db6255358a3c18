"""Streaming file-level operators: channelize / synthesize DADA files of any length through the GPU
without holding the payload in host or device memory.

The reference moves data between stages as whole files read into memory
(matlab/DADARead.m:58-83 reads a block of samples at an offset; python/data_gen/pipeline.py:62-84
chains file -> array -> file).  Here the TFP payload is cut on block boundaries, each piece is read
straight into pinned host memory by a reader thread, copied to the device, transformed and copied back
on three CUDA streams, and written at its final offset of the output file by a writer thread, with
``depth`` pieces in flight.  Pieces carry their own FIR history / overlap-save overlap (re-read from the
file), so the result is bit-identical to the whole-file operators of ``format_handler``.

Integer payloads (NBIT 8 / 16, DADARead.m:41-53) travel as integers - disk, pinned memory and PCIe carry 2-4x fewer
bytes - and are cast to the arithmetic type on the device (``pfb_unpack_payload``), bit-identically to a host cast.

The padded (SKA-Mid) analysis wraps its output circularly (polyphase_analysis_padded.m:156) and is not
streamable; ``stream_channelize`` refuses it.
"""
from __future__ import annotations

import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from .formats import DADAFile, format_header, read_header
from .rational import Rational


def _payload_info(path):
    """(header, arithmetic complex dtype, payload offset, samples, NBIT).  NBIT 32 / 64 payloads are complex floats;
    NBIT 8 / 16 are (re, im) integer pairs (DADARead.m:41-53) that travel as integers and are cast on the device."""
    hdr = read_header(path)
    nbit = int(hdr["NBIT"])
    if int(hdr.get("NDIM", "2")) != 2 or nbit not in (8, 16, 32, 64):
        raise ValueError("streaming operators need a complex payload (NDIM 2) with NBIT 8, 16, 32 or 64")
    cdt = np.complex128 if nbit == 64 else np.complex64
    off = int(hdr["HDR_SIZE"])
    per = int(hdr["NCHAN"]) * int(hdr["NPOL"])
    ndat = (os.path.getsize(path) - off) // (per * 2 * nbit // 8)
    return hdr, cdt, off, ndat, nbit


def _int_input(torch, nbit):
    """(torch dtype of the travelling payload, elements per complex sample) for an NBIT 8 / 16 payload, else None."""
    return {8: (torch.int8, 2), 16: (torch.int16, 2)}.get(nbit)


def _unpacker(torch, device, nbit, max_complex, tdtype):
    """Device-side int -> complex cast (pfb_unpack_payload) into one staging buffer owned by the pipeline."""
    import ctypes
    from . import lib as _l
    stage = torch.empty((max_complex,), dtype=tdtype, device=device)
    code = _l.PFB_DTYPE_C64 if tdtype == torch.complex64 else _l.PFB_DTYPE_C128

    def unpack(din_int, ncomplex):
        st = ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
        _l.check(_l.lib.pfb_unpack_payload(din_int.data_ptr(), stage.data_ptr(), 2 * ncomplex, nbit, code, st))
        return stage[:ncomplex]
    return unpack


class _Pipe:
    """depth-deep ring of (pinned in, device in, device out, pinned out) buffers + the three streams.

    Every buffer is guarded by the event (or future) of ITS last user only, so `depth` pieces are genuinely in
    flight: disk read of piece i+depth-1, H2D of piece i+1, kernels of piece i, D2H of piece i-1 and the disk write
    of piece i-2 overlap."""

    def __init__(self, torch, device, depth, in_shape, out_shape, tdtype, in_dtype=None):
        self.torch, self.depth = torch, depth
        in_dtype = in_dtype or tdtype
        self.hin = [torch.empty(in_shape, dtype=in_dtype).pin_memory() for _ in range(depth)]
        self.hout = [torch.empty(out_shape, dtype=tdtype).pin_memory() for _ in range(depth)]
        self.din = [torch.empty(in_shape, dtype=in_dtype, device=device) for _ in range(depth)]
        self.dout = [torch.empty(out_shape, dtype=tdtype, device=device) for _ in range(depth)]
        self.s_in, self.s_cmp, self.s_out = (torch.cuda.Stream(device) for _ in range(3))
        self.ev_h2d = [None] * depth         # H2D out of hin[slot] finished  -> the reader may refill hin[slot]
        self.ev_free = [None] * depth        # kernels finished with din[slot] -> the next H2D may overwrite it
        self.ev_done = [None] * depth        # D2H out of dout[slot] finished  -> the next kernels may overwrite it
        self.wr = [None] * depth             # writer future still reading hout[slot]
        self.reader = ThreadPoolExecutor(max_workers=1)
        self.writer = ThreadPoolExecutor(max_workers=1)

    def close(self):
        self.reader.shutdown(wait=True)
        self.writer.shutdown(wait=True)


def _read_into(path, offset, dst_numpy, event):
    if event is not None:
        event.synchronize()                  # the H2D copy that last read this pinned buffer
    with open(path, "rb", buffering=0) as f:
        f.seek(offset)
        view = memoryview(dst_numpy).cast("B")
        got = 0
        while got < len(view):
            n = f.readinto(view[got:])
            if not n:
                raise IOError(f"{path}: short read at byte {offset + got}")
            got += n


def _write_from(fd, offset, src_numpy, event):
    event.synchronize()
    view = memoryview(src_numpy).cast("B")
    done = 0
    while done < len(view):                  # pwrite may write less than asked
        done += os.pwrite(fd, view[done:], offset + done)


def _run(pipe, pieces, in_path, in_item, fd_out, out_item, compute):
    """pieces: list of (in_byte_offset, in_elems, out_byte_offset, out_elems, args).  compute(din_view, dout_view, args)
    enqueues the kernels on the current stream."""
    torch = pipe.torch
    reads = {}

    def start_read(i):
        slot = i % pipe.depth
        ib, ie = pieces[i][0], pieces[i][1]
        dst = pipe.hin[slot].view(-1)[:ie].numpy()
        reads[i] = pipe.reader.submit(_read_into, in_path, ib, dst, pipe.ev_h2d[slot])

    for i in range(min(pipe.depth, len(pieces))):
        start_read(i)
    for i, (ib, ie, ob, oe, args) in enumerate(pieces):
        slot = i % pipe.depth
        reads.pop(i).result()                               # piece i is in pinned memory
        with torch.cuda.stream(pipe.s_in):
            if pipe.ev_free[slot] is not None:
                pipe.s_in.wait_event(pipe.ev_free[slot])
            pipe.din[slot].view(-1)[:ie].copy_(pipe.hin[slot].view(-1)[:ie], non_blocking=True)
            pipe.ev_h2d[slot] = torch.cuda.Event()
            pipe.ev_h2d[slot].record(pipe.s_in)
        if i + pipe.depth < len(pieces):
            start_read(i + pipe.depth)                      # refills hin[slot] once the copy above has drained it
        with torch.cuda.stream(pipe.s_cmp):
            pipe.s_cmp.wait_event(pipe.ev_h2d[slot])
            if pipe.ev_done[slot] is not None:
                pipe.s_cmp.wait_event(pipe.ev_done[slot])
            compute(pipe.din[slot].view(-1)[:ie], pipe.dout[slot].view(-1)[:oe], args)
            pipe.ev_free[slot] = torch.cuda.Event()
            pipe.ev_free[slot].record(pipe.s_cmp)
        if pipe.wr[slot] is not None:
            pipe.wr[slot].result()                          # hout[slot]: the write of piece i - depth (long finished)
        with torch.cuda.stream(pipe.s_out):
            pipe.s_out.wait_event(pipe.ev_free[slot])
            pipe.hout[slot].view(-1)[:oe].copy_(pipe.dout[slot].view(-1)[:oe], non_blocking=True)
            pipe.ev_done[slot] = torch.cuda.Event()
            pipe.ev_done[slot].record(pipe.s_out)
        pipe.wr[slot] = pipe.writer.submit(_write_from, fd_out, ob, pipe.hout[slot].view(-1)[:oe].numpy(), pipe.ev_done[slot])
    for w in pipe.wr:
        if w is not None:
            w.result()


def stream_channelize(channelizer, input_file_path, output_dir="./", output_file_name=None, blocks_per_piece=4096,
                      depth=3):
    """``format_handler.PSRFormatChannelizer`` on a file of any length.  Returns the output DADAFile (header set,
    payload on disk; ``load_data()`` reads it)."""
    from .format_handler import _fmt_num, _tdtype, _torch
    torch = _torch()
    hdr, cdt, off, ndat, nbit = _payload_info(input_file_path)
    intin = _int_input(torch, nbit)
    if int(hdr["NCHAN"]) != 1:
        raise ValueError("expected single-channel input")
    npol = int(hdr["NPOL"])
    plan = channelizer._plan(_tdtype(cdt))
    if plan.variant == 1:
        raise ValueError("the padded analysis wraps its output circularly and cannot be streamed")
    first_call = True
    # pieces start on multiples of 16 blocks: the shape-specialised kernels then see the same block alignment as in a
    # whole-file run and every block is computed by the same kernel, i.e. bit-identically
    blocks_per_piece = max(16, (int(blocks_per_piece) + 15) // 16 * 16)
    nblocks = plan.nblocks(ndat, first_call)
    S, fl, nco = plan.step, plan.fl, plan.nchan_out
    pre = 1536 if plan.variant == 2 else 0                  # lowcbf pre-pad (PSTFilterbank.m:27-34)
    item = np.dtype(cdt).itemsize                                 # output (and arithmetic) sample size
    in_item, in_per = (2 * nbit // 8, intin[1]) if intin else (item, 1)   # bytes / travelling elements per input sample
    if output_file_name is None:
        output_file_name = "channelized." + os.path.basename(input_file_path)
    of = DADAFile(os.path.join(output_dir, output_file_name), header=hdr)
    of["TSAMP"] = _fmt_num(float(hdr.get("TSAMP", "1")) * channelizer.nchan * channelizer.os_factor.de / channelizer.os_factor.nu)
    of["PFB_DC_CHAN"] = "1"
    of["NCHAN_PFB_0"] = str(channelizer.nchan)
    of["OS_FACTOR"] = str(channelizer.os_factor)
    of["NSTAGE"] = "1"
    of["COEFF_0"] = ",".join("%0.6E" % c for c in channelizer.fir_filter_coeff)
    of["OVERSAMP_0"] = str(channelizer.os_factor)
    of["NTAP_0"] = str(channelizer.fir_filter_coeff.size)
    of["NCHAN"], of["NPOL"], of["NBIT"], of["NDIM"] = str(nco), str(npol), str(item * 4), "2"
    head = format_header(of.header)
    os.makedirs(os.path.dirname(os.path.abspath(of.file_path)), exist_ok=True)
    pieces = []
    for k0 in range(0, nblocks, blocks_per_piece):
        k1 = min(nblocks, k0 + blocks_per_piece)
        s0 = max(0, k0 * S - pre)
        s1 = min(ndat, (k1 - 1) * S - pre + fl)
        pieces.append((off + s0 * npol * in_item, (s1 - s0) * npol * in_per, len(head) + k0 * nco * npol * item,
                       (k1 - k0) * nco * npol, (s0, k0, k1 - k0, s1 - s0)))
    max_in = max((p[1] for p in pieces), default=0)
    max_out = max((p[3] for p in pieces), default=0)
    fd = os.open(of.file_path, os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o644)
    try:
        os.pwrite(fd, head, 0)
        if pieces:
            pipe = _Pipe(torch, plan.device, depth, (max_in,), (max_out,), _tdtype(cdt), intin[0] if intin else None)
            unpack = _unpacker(torch, plan.device, nbit, max_in // in_per, _tdtype(cdt)) if intin else None

            def compute(din, dout, args):
                s0, k0, nb, ns = args
                if unpack is not None:
                    din = unpack(din, ns * npol)
                plan.segment(din.view(ns, npol), s0, k0, nb, out=dout.view(nb, nco, npol), first_call=first_call)

            try:
                _run(pipe, pieces, input_file_path, in_item, fd, item, compute)
            finally:
                pipe.close()
    finally:
        os.close(fd)
    of.header = read_header(of.file_path)
    of._lazy = True                       # .data maps the payload on first access (no copy into memory here)
    return of


def stream_synthesize(synthesizer, input_file_path, output_dir="./", output_file_name=None, blocks_per_piece=64,
                      depth=3):
    """``format_handler.PSRFormatSynthesizer`` on a file of any length."""
    from .format_handler import _fmt_num, _tdtype, _torch
    torch = _torch()
    hdr, cdt, off, ndat, nbit = _payload_info(input_file_path)
    intin = _int_input(torch, nbit)
    nchan, npol = int(hdr["NCHAN"]), int(hdr["NPOL"])
    os_factor = Rational.from_str(hdr["OS_FACTOR"])
    taps = None
    if synthesizer.apply_deripple:
        if "COEFF_0" not in hdr:
            raise ValueError("deripple requested but no filter coefficients available (COEFF_0 missing)")
        taps = np.array([float(v) for v in hdr["COEFF_0"].split(",")])
    plan = synthesizer._plan(nchan, os_factor, taps, _tdtype(cdt))
    nblocks = plan.nblocks(ndat)
    keep, nf = synthesizer.input_fft_length - 2 * synthesizer.input_overlap, synthesizer.input_fft_length
    okeep = plan.out_len(keep + 2 * synthesizer.input_overlap)        # output samples per block
    item = np.dtype(cdt).itemsize
    in_item, in_per = (2 * nbit // 8, intin[1]) if intin else (item, 1)
    if output_file_name is None:
        output_file_name = "synthesized." + os.path.basename(input_file_path)
    of = DADAFile(os.path.join(output_dir, output_file_name), header=hdr)
    of["TSAMP"] = _fmt_num(float(hdr.get("TSAMP", "1")) / (os_factor.de / os_factor.nu) / nchan)       # synthesize.m:92
    of["NCHAN"], of["NPOL"], of["NBIT"], of["NDIM"] = "1", str(npol), str(item * 4), "2"
    head = format_header(of.header)
    os.makedirs(os.path.dirname(os.path.abspath(of.file_path)), exist_ok=True)
    row = nchan * npol
    pieces = []
    for n0 in range(0, nblocks, blocks_per_piece):
        n1 = min(nblocks, n0 + blocks_per_piece)
        rows = (n1 - n0 - 1) * keep + nf
        pieces.append((off + n0 * keep * row * in_item, rows * row * in_per, len(head) + n0 * okeep * npol * item,
                       (n1 - n0) * okeep * npol, (rows, n1 - n0)))
    max_in = max((p[1] for p in pieces), default=0)
    max_out = max((p[3] for p in pieces), default=0)
    fd = os.open(of.file_path, os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o644)
    try:
        os.pwrite(fd, head, 0)
        if pieces:
            pipe = _Pipe(torch, plan.device, depth, (max_in,), (max_out,), _tdtype(cdt), intin[0] if intin else None)
            unpack = _unpacker(torch, plan.device, nbit, max_in // in_per, _tdtype(cdt)) if intin else None

            def compute(din, dout, args):
                rows, nb = args
                if unpack is not None:
                    din = unpack(din, rows * row)
                plan(din.view(rows, nchan, npol), out=dout.view(nb * okeep, npol))

            try:
                _run(pipe, pieces, input_file_path, in_item, fd, item, compute)
            finally:
                pipe.close()
    finally:
        os.close(fd)
    of.header = read_header(of.file_path)
    of._lazy = True                       # .data maps the payload on first access (no copy into memory here)
    return of
