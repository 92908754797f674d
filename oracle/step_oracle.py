"""Pure-Python restatement of SPEC.md §2-§4 (one session) — TEST INFRASTRUCTURE.

Small cases only; used to cross-check oracle/abr_oracle.c and as the timed
"reference-style" CPU port of the chunk step (the reference's own step loop,
Simulator.py:93-210, does not run: SURVEY.md D1-D6).  Parity unpinned by the
reference; SPEC.md is the contract.  Also holds ``euler_session`` — the
intended fixed-dt dynamics of Simulator.py:135-208 restated for a loose
plausibility check (|Δdelay| ≲ dt per chunk) against the analytic walk.
"""
from __future__ import annotations

import math


class Session:
    def __init__(self, bw, interval, sizes, util, P, start_offset=0.0):
        self.bw, self.I, self.T = list(bw), float(interval), len(bw)
        self.sizes, self.util, self.P = sizes, util, P
        self.V, self.A = len(sizes), len(sizes[0])
        n = math.floor(start_offset / self.I)
        self.seg = int(math.fmod(n, self.T))
        self.tau = start_offset - n * self.I
        if self.tau < 0.0:
            self.tau = 0.0
        if self.tau >= self.I:
            self.tau = 0.0
            self.seg = (self.seg + 1) % self.T
        self.buffer = 0.0
        self.chunk = 0
        self.last_q = P["default_quality"]
        self.done = False
        self.history = []          # all measured throughputs since reset

    def step(self, q):
        P = self.P
        if self.done:
            return dict(delay=0.0, sleep=0.0, buffer=self.buffer, rebuf=0.0, reward=0.0, eov=1, inert=True)
        size = self.sizes[self.chunk][q]
        sent = 0.0
        k = 0
        room0 = self.I - self.tau
        rate = self.bw[self.seg] * P["payload"]
        cap = rate * room0
        while True:                                  # SPEC 3.1 (Simulator.py:158-163)
            if sent + cap >= size:
                break
            sent = sent + cap
            k += 1
            self.seg = 0 if self.seg + 1 == self.T else self.seg + 1
            self.tau = 0.0
            rate = self.bw[self.seg] * P["payload"]
            cap = rate * self.I
        dt = (size - sent) / rate
        self.tau = self.tau + dt
        elapsed = 0.0 if k == 0 else room0 + float(k - 1) * self.I
        delay = (elapsed + dt) + P["rtt"]
        thr = size / delay
        rebuf = delay - self.buffer if delay - self.buffer > 0 else 0.0     # SPEC 3.2
        self.buffer = (self.buffer - delay if self.buffer - delay > 0 else 0.0) + P["chunk_length"]
        sleep = 0.0
        if self.buffer > P["max_buffer"]:                                     # SPEC 3.3
            sleep = math.ceil((self.buffer - P["max_buffer"]) / P["sleep_quantum"]) * P["sleep_quantum"]
            self.buffer = self.buffer - sleep
            x = self.tau + sleep
            n = math.floor(x / self.I)
            self.tau = x - n * self.I
            self.seg = (self.seg + n) % self.T
            if self.tau < 0.0:
                self.tau = 0.0
            if self.tau >= self.I:
                self.tau = 0.0
                self.seg = (self.seg + 1) % self.T
        u = self.util[self.chunk][q]                                          # SPEC 3.4
        smooth = abs(u - self.util[self.chunk][self.last_q]) if self.last_q >= 0 else 0.0
        reward = (u - P["rebuf_penalty"] * rebuf) - P["smooth_penalty"] * smooth
        self.history.append(thr)
        self.last_q = q                                                       # SPEC 3.5
        self.chunk += 1
        eov = self.chunk >= self.V
        out = dict(delay=delay, sleep=sleep, buffer=self.buffer, rebuf=rebuf, reward=reward, eov=int(eov),
                   throughput=thr, u=u, smooth=smooth, inert=False)
        if eov and P["auto_reset"]:
            self.chunk = 0
            self.buffer = 0.0
            self.last_q = P["default_quality"]
            self.history = []
        elif eov:
            self.done = True
        return out


def bba_action(buffer, A, reservoir, cushion):
    if buffer < reservoir:
        return 0
    if buffer >= reservoir + cushion:
        return A - 1
    return min(A - 1, int(math.floor(((A - 1) * (buffer - reservoir)) / cushion)))


def euler_download_delay(bw, interval, start_time, size, payload, dt=0.01):
    """Intended download block of Simulator.py:152-170: accumulate bw·dt per tick until
    downloaded_size >= target_size; returns download_time (no RTT).  Loose check only."""
    t = start_time
    got = 0.0
    elapsed = 0.0
    T = len(bw)
    while got < size:
        b = bw[int(t / interval) % T] * payload
        got += b * dt
        elapsed += dt
        t += dt
    return elapsed
