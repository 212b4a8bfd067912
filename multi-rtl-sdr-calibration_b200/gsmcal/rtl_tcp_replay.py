"""Replay server speaking the rtl_tcp wire protocol - test / bring-up infrastructure.

Stands in for `rtl_tcp -p <port> -d <i>` (gsm_sync_demod.m:4-8): on connect it sends the 12-byte greeting
("RTL0", tuner type, gain count; big-endian), then streams raw interleaved uint8 I,Q of a stored capture (cyclically),
while accepting the 5-byte commands the reference sends (set_freq_tcp.m, set_rate_tcp.m, set_gain_tcp.m) and
recording them.  With `rate_bytes_per_s` set the stream is paced like a real dongle (2 bytes x sample rate), else it
runs as fast as the socket takes it.  With it the untouched .m drivers (or gsmcal.ingest) run end to end without a dongle.

  python -m gsmcal.rtl_tcp_replay --port 1234 --dongles 2 --seconds 0.5        # synthetic GSM C0 captures
"""
from __future__ import annotations

import socket
import struct
import threading
import time

import numpy as np

GREETING = b"RTL0" + struct.pack(">II", 5, 29)      # R820T, 29 gain steps


class ReplayDongle:
    def __init__(self, capture_u8: np.ndarray, port: int = 0, host: str = "127.0.0.1",
                 rate_bytes_per_s: float | None = None, chunk: int = 1 << 18):
        self.capture = np.ascontiguousarray(capture_u8, dtype=np.uint8).ravel()
        self.rate = rate_bytes_per_s
        self.chunk = int(chunk)
        self.commands: list[tuple[int, int]] = []
        self.bytes_sent = 0
        self._srv = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
        self._srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        self._srv.bind((host, port))
        self._srv.listen(1)
        self.host, self.port = self._srv.getsockname()
        self._stop = threading.Event()
        self._thread = threading.Thread(target=self._serve, daemon=True)
        self._thread.start()

    # one client at a time, like rtl_tcp
    def _serve(self):
        self._srv.settimeout(0.2)
        while not self._stop.is_set():
            try:
                conn, _ = self._srv.accept()
            except socket.timeout:
                continue
            except OSError:
                return
            with conn:
                conn.setsockopt(socket.SOL_SOCKET, socket.SO_SNDBUF, 4 << 20)
                rx = threading.Thread(target=self._commands, args=(conn,), daemon=True)
                rx.start()
                try:
                    self._stream(conn)
                except (BrokenPipeError, ConnectionResetError, OSError):
                    pass

    def _commands(self, conn: socket.socket):
        buf = b""
        try:
            while not self._stop.is_set():
                d = conn.recv(4096)
                if not d:
                    return
                buf += d
                while len(buf) >= 5:
                    cmd, param = struct.unpack(">BI", buf[:5])
                    self.commands.append((cmd, param))
                    buf = buf[5:]
        except OSError:
            return

    def _stream(self, conn: socket.socket):
        conn.sendall(GREETING)
        mv = memoryview(self.capture)
        n, off, t0, sent = len(mv), 0, time.perf_counter(), 0
        while not self._stop.is_set():
            m = min(self.chunk, n - off)
            conn.sendall(mv[off:off + m])
            off = (off + m) % n
            sent += m
            self.bytes_sent += m
            if self.rate:
                ahead = sent / self.rate - (time.perf_counter() - t0)
                if ahead > 0:
                    time.sleep(ahead)

    def expected(self, first_byte: int, n_bytes: int) -> np.ndarray:
        """Bytes [first_byte, first_byte+n_bytes) of the IQ stream that follows the greeting (cyclic capture)."""
        idx = (first_byte + np.arange(n_bytes, dtype=np.int64)) % len(self.capture)
        return self.capture[idx]

    def close(self):
        self._stop.set()
        try:
            self._srv.close()
        except OSError:
            pass
        self._thread.join(timeout=2)


def main():
    import argparse
    import torch
    from . import synth
    ap = argparse.ArgumentParser()
    ap.add_argument("--port", type=int, default=1234)
    ap.add_argument("--dongles", type=int, default=2)
    ap.add_argument("--seconds", type=float, default=0.5)
    ap.add_argument("--unpaced", action="store_true")
    a = ap.parse_args()
    fs = (1625.0 / 6.0) * 1e3 * 8
    n = int(a.seconds * fs)
    specs = [synth.random_spec(seed, n) for seed in range(1, a.dongles + 1)]
    raw = synth.generate_batch(specs, device="cuda" if torch.cuda.is_available() else "cpu")
    servers = [ReplayDongle(raw[d].cpu().numpy(), a.port + d, rate_bytes_per_s=None if a.unpaced else 2 * fs)
               for d in range(a.dongles)]
    print("replaying on ports", [s.port for s in servers], flush=True)
    try:
        while True:
            time.sleep(1)
    except KeyboardInterrupt:
        for s in servers:
            s.close()


if __name__ == "__main__":
    main()
