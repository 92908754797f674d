"""Pure-Python restatement of SPEC.md §2-§4 (one session) — TEST INFRASTRUCTURE.

Small cases only; used to cross-check oracle/abr_oracle.c and as the timed
"reference-style" CPU port of the chunk step (the reference's own step loop,
Simulator.py:93-210, does not run as shipped: SURVEY.md D1-D6).  Parity: the live-mode
dynamics (SPEC §7) are pinned by that loop itself, repaired in memory and executed
(oracle/make_ref_simulator.py -> tests/golden/sim_ref_tick_golden.json); the north-star
constants (RTT, payload factor, sleep quantum, per-step reward) are defined by SPEC.md.  ``Session(..., walk="segments")`` replaces the
cumulative-capacity form of SPEC §3.1 by the segment-by-segment integration it is the
closed form of (the running sum restarts at the session's position instead of the start
of the trace), which tests use to bound the difference (≪ 1e-9 relative).  Also holds ``euler_session`` — the
intended fixed-dt dynamics of Simulator.py:135-208 restated for a loose
plausibility check (|Δdelay| ≲ dt per chunk) against the analytic walk.
"""
from __future__ import annotations

import math


def capacity_table(bw, interval, payload):
    """SPEC §3.1: (rate, C) with C[0] = 0, C[j+1] = C[j] + (bw[j]*payload)*interval, left to right.
    ``rate`` is used by walk="segments" only."""
    rate = [b * payload for b in bw]
    C = [0.0]
    for r in rate:
        C.append(C[-1] + r * interval)
    return rate, C


class Session:
    def __init__(self, bw, interval, sizes, util, P, start_offset=0.0, walk="table", table=None):
        self.bw, self.I, self.T = list(bw), float(interval), len(bw)
        self.sizes, self.util, self.P = sizes, util, P
        self.V, self.A = len(sizes), len(sizes[0])
        self.walk = walk
        # SPEC 3.1 table (per trace; pass ``table=capacity_table(...)`` to share it between sessions of a trace)
        self.rate, self.C = table if table is not None else capacity_table(self.bw, self.I, P["payload"])
        x = start_offset / self.I
        n = math.floor(x)
        self.seg = int(math.fmod(n, self.T))
        self.phi = x - n                      # fraction of segment seg already consumed (SPEC §1)
        self.pos = self._position()           # the same position in data coordinates (SPEC §3.1)
        self.buffer = 0.0
        self.chunk = 0
        self.last_q = P["default_quality"]
        self.done = False
        self.history = []          # all measured throughputs since reset
        # live mode (SPEC §7)
        self.t_now = 0.0
        self.play_time = 0.0
        self.play_id = 0           # content chunk being played, and how much of it has been played
        self.play_len = 0.0
        self.started = P.get("start_up_length", 0.0) <= 0.0
        self.speed = None          # playback speed per content chunk (list of V), None = 1.0

    def _speed_of(self, k):
        if self.speed is None:
            return 1.0
        return float(self.speed[k if k < self.V else self.V - 1])

    def _piece(self, d, dw, v, acc):
        """One stretch of playback: d seconds of content in dw seconds of wall time at speed v."""
        # integral of the latency (wall clock - content played) over the stretch; it changes at the rate 1 - v
        acc["area"] = acc["area"] + ((acc["tc"] - self.play_time) * dw + ((1.0 - v) * dw) * (dw * 0.5))
        self.play_time = self.play_time + d
        self.buffer = self.buffer - d
        acc["tc"] = acc["tc"] + dw
        acc["played"] = acc["played"] + d

    def _play_wall(self, dt, acc):
        """SPEC §7 play_wall(Δ): playback during Δ seconds of wall time; returns the stall time."""
        L = self.P["chunk_length"]
        if not self.started:
            acc["startup"] = acc["startup"] + dt
            acc["tc"] = acc["tc"] + dt
            return 0.0
        rem = dt
        while rem > 0.0 and self.buffer > 0.0:
            v = self._speed_of(self.play_id)
            room = L - self.play_len
            can = room if room < self.buffer else self.buffer
            need = v * rem
            if need < can:
                self._piece(need, rem, v, acc)
                self.play_len = self.play_len + need
                rem = 0.0
            else:
                dw = can / v
                finished = room <= self.buffer          # the chunk ends before the buffer does
                self._piece(can, dw, v, acc)
                rem = rem - dw
                if finished:
                    self.play_id += 1
                    self.play_len = 0.0
                else:
                    self.play_len = self.play_len + can
        if rem < 0.0:
            rem = 0.0
        acc["tc"] = acc["tc"] + rem
        return rem

    def _play_content(self, x, acc):
        """SPEC §7 play_content(X): playback until X seconds of content have drained; returns the wall time."""
        L = self.P["chunk_length"]
        w = 0.0
        while x > 0.0:
            v = self._speed_of(self.play_id)
            room = L - self.play_len
            if room <= x:
                d, finished = room, True
            else:
                d, finished = x, False
            dw = d / v
            self._piece(d, dw, v, acc)
            w = w + dw
            if finished:
                self.play_id += 1
                self.play_len = 0.0
                x = x - d
            else:
                self.play_len = self.play_len + d
                x = 0.0
        return w

    def _position(self):
        C = self.C
        return C[self.seg] + (C[self.seg + 1] - C[self.seg]) * self.phi

    def _advance(self, dt):
        x = self.phi + dt / self.I
        n = math.floor(x)
        self.phi = x - n
        self.seg = (self.seg + n) % self.T
        self.pos = self._position()

    def _download_segments(self, size):
        """Segment-by-segment integration from the current position (not the SPEC's arithmetic: cross-check only)."""
        sent = 0.0
        k = 0
        tau = self.phi * self.I
        room0 = self.I - tau
        rate = self.rate[self.seg]
        cap = rate * room0
        while True:
            if sent + cap >= size:
                break
            sent = sent + cap
            k += 1
            self.seg = 0 if self.seg + 1 == self.T else self.seg + 1
            tau = 0.0
            rate = self.rate[self.seg]
            cap = rate * self.I
        dt = (size - sent) / rate
        self.phi = (tau + dt) / self.I
        return (0.0 if k == 0 else room0 + float(k - 1) * self.I) + dt

    def step(self, q, speed=None):
        """One chunk step.  ``speed`` (live mode): playback speed per content chunk, a list of V values (kept for the
        following steps), or None to keep the table set before (1.0 initially)."""
        P = self.P
        if speed is not None:
            self.speed = list(speed)
        if self.done:
            return dict(delay=0.0, sleep=0.0, buffer=self.buffer, rebuf=0.0, reward=0.0, eov=1, inert=True,
                        latency=0.0, startup=0.0, area=0.0, played=0.0)
        size = self.sizes[self.chunk][q]
        live = bool(P.get("live", 0))
        acc = dict(startup=0.0, area=0.0, played=0.0, tc=self.t_now)
        idle = 0.0
        rebuf = 0.0
        if live:                                                              # SPEC 7.1
            w1 = (self.chunk + 1) * P["chunk_length"] - self.t_now
            w1 = w1 if w1 > 0 else 0.0
            rebuf = self._play_wall(w1, acc)
            w2 = self._play_content(self.buffer - P["max_buffer"], acc) \
                if (self.started and self.buffer > P["max_buffer"]) else 0.0
            idle = w1 + w2
            if idle > 0:
                self._advance(idle)
        if self.walk == "segments":
            delay = self._download_segments(size) + P["rtt"]
        else:                                        # SPEC 3.1 (Simulator.py:158-163 in closed form)
            C, T = self.C, self.T
            target = self.pos + size
            n = 0
            while target >= C[T]:
                target = target - C[T]
                n += 1
            j = self.seg if n == 0 else 0          # C[seg] <= pos <= target without a wrap
            while j + 1 < T and C[j + 1] <= target:
                j += 1
            phi_new = (target - C[j]) / (C[j + 1] - C[j])
            k = (j - self.seg) + n * T
            dl = (float(k) + (phi_new - self.phi)) * self.I
            delay = (dl if dl > 0 else 0.0) + P["rtt"]
            self.seg, self.phi, self.pos = j, phi_new, target
        thr = size / delay
        latency = 0.0
        if live:                                                              # SPEC 7.2
            rebuf = rebuf + self._play_wall(delay, acc)
            self.buffer = self.buffer + P["chunk_length"]
            self.t_now = (self.t_now + idle) + delay
            if not self.started and self.buffer >= P.get("start_up_length", 0.0):
                self.started = True
            latency = self.t_now - self.play_time
            sleep = idle
        else:
            rebuf = delay - self.buffer if delay - self.buffer > 0 else 0.0     # SPEC 3.2
            self.buffer = (self.buffer - delay if self.buffer - delay > 0 else 0.0) + P["chunk_length"]
            sleep = 0.0
        if not live and self.buffer > P["max_buffer"]:                        # SPEC 3.3
            sleep = math.ceil((self.buffer - P["max_buffer"]) / P["sleep_quantum"]) * P["sleep_quantum"]
            self.buffer = self.buffer - sleep
            self._advance(sleep)
        u = self.util[self.chunk][q]                                          # SPEC 3.4
        # the previous index is looked up in the current chunk's ladder (mpc.py:148-149) or, smooth_prev_ladder = 1, in
        # the previous chunk's own ladder (Simulator.calculate_qoe, Simulator.py:81-82)
        prow = self.chunk - 1 if (P.get("smooth_prev_ladder", 0) and self.chunk > 0) else self.chunk
        smooth = abs(u - self.util[prow][self.last_q]) if self.last_q >= 0 else 0.0
        reward = (u - P["rebuf_penalty"] * rebuf) - P["smooth_penalty"] * smooth
        if live:
            reward = reward - P.get("latency_penalty", 0.0) * latency
        self.history.append(thr)
        self.last_q = q                                                       # SPEC 3.5
        self.chunk += 1
        eov = self.chunk >= self.V
        out = dict(delay=delay, sleep=sleep, buffer=self.buffer, rebuf=rebuf, reward=reward, eov=int(eov),
                   throughput=thr, u=u, smooth=smooth, inert=False, latency=latency, startup=acc["startup"],
                   area=acc["area"], played=acc["played"])
        if eov and P["auto_reset"]:
            self.chunk = 0
            self.buffer = 0.0
            self.last_q = P["default_quality"]
            self.history = []
            self.t_now = 0.0
            self.play_time = 0.0
            self.play_id = 0
            self.play_len = 0.0
            self.started = P.get("start_up_length", 0.0) <= 0.0
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


def euler_live_session(bw, interval, sizes, L, B, start_up_length, payload, speed=1.0, dt=0.001):
    """The intended live dynamics of Simulator.py:135-208 as a fixed-dt loop, written from the tick order in
    SURVEY.md §3.2 with the two missing pause resets supplied (D2): timers, live-edge/buffer-full gate, download
    against the square wave, playback drain at `speed`, start-up latch.  `sizes[k]` is the payload of chunk k.
    Returns dict(rebuffer, startup, latency, delays).  Loose plausibility check for SPEC §7 only."""
    T = len(bw)
    t = 0.0
    chunk = 0
    buffer = 0.0
    start_up = start_up_length > 0
    rebuffer = startup = play_time = 0.0
    got = dtime = 0.0
    delays = []
    while chunk < len(sizes):
        if start_up:
            startup += dt
        elif buffer <= 0.0:
            rebuffer += dt
        available = int(t / L) - 1
        if available >= chunk and (buffer < B or dtime > 0.0):
            got += bw[int(t / interval) % T] * payload * dt
            dtime += dt
            if got >= sizes[chunk]:
                delays.append(dtime)
                chunk += 1
                got = dtime = 0.0
                buffer += L
        if not start_up and buffer > 0.0:
            play_time += speed * dt
            buffer -= speed * dt
        if buffer <= 0.0:
            buffer = 0.0
        if start_up and buffer >= start_up_length:
            start_up = False
        t += dt
    return dict(rebuffer=rebuffer, startup=startup, latency=t - play_time, delays=delays, t=t, play_time=play_time)
