"""Discrete-event model of the scan kernel's ready-queue schedule (csrc/dm_mamba1.cu, kDyn = true), checked on CPU.

The kernel's persistent warps take consumer tickets from an atomic head; tickets below n_units are the first segments
(ready from the start), ticket n_units + j is whatever the j-th hand-over pushes (atomic tail), and a warp that finishes
segment s of a unit pushes (unit, s + 1) unless s was the last.  The claims the kernel's comment makes are verified
here for random item durations and ANY number of workers, including fewer workers than units and a single worker:

* every (unit, segment) runs exactly once and segment s of a unit starts only after segment s - 1 finished
  (items never wait once started: their input state exists);
* no worker ever polls a queue slot that is never filled (no deadlock) and all workers terminate;
* the number of pushes equals n_items - n_units, so the queue capacity the C-ABI sizes (64 segments per unit) holds.
"""
import heapq
import random

import pytest


def simulate(n_units, n_segs, n_workers, rng):
    n_items = n_units * n_segs
    head = 0                      # consumer tickets handed out
    queue = []                    # pushes in completion order: item ids
    done_at = {}                  # item -> finish time
    started_at = {}
    waiting = {}                  # queue slot index -> worker polling it
    events = []                   # (time, worker, item) completion events
    free = list(range(n_workers))
    t = 0.0
    exited = 0

    def take(worker, now):
        """worker asks for its next item at time `now`; returns True if it got one or exited, False if it polls."""
        nonlocal head, exited
        ticket = head
        head += 1
        if ticket >= n_items:
            exited += 1
            return True
        if ticket < n_units:
            start(worker, ticket, now)
            return True
        slot = ticket - n_units
        if slot < len(queue):
            start(worker, queue[slot], now)
            return True
        assert slot not in waiting
        waiting[slot] = worker
        return False

    def start(worker, item, now):
        assert item not in started_at, "item handed out twice"
        seg, unit = divmod(item, n_units)
        if seg > 0:
            prev = (seg - 1) * n_units + unit
            assert prev in done_at and done_at[prev] <= now, "segment started before its predecessor finished"
        started_at[item] = now
        heapq.heappush(events, (now + rng.uniform(0.5, 1.5), worker, item))

    for w in free:
        take(w, 0.0)
    while events:
        t, w, item = heapq.heappop(events)
        done_at[item] = t
        seg, unit = divmod(item, n_units)
        if seg + 1 < n_segs:                       # hand-over: publish the successor
            queue.append((seg + 1) * n_units + unit)
            slot = len(queue) - 1
            if slot in waiting:
                start(waiting.pop(slot), queue[slot], t)
        take(w, t)
    assert not waiting, f"workers left polling slots {sorted(waiting)} that are never filled"
    assert exited == n_workers
    assert len(done_at) == n_items and len(queue) == n_items - n_units
    return t


@pytest.mark.parametrize("n_units,n_segs,n_workers", [(1536, 5, 1776), (3072, 3, 1776), (7, 4, 3), (5, 6, 1), (40, 2, 64),
                                                      (16, 1, 4), (3, 64, 2)])
def test_ready_queue_schedule_runs_every_item_once_and_terminates(n_units, n_segs, n_workers):
    rng = random.Random(n_units * 131 + n_segs * 17 + n_workers)
    simulate(n_units, n_segs, n_workers, rng)


def test_ready_queue_balances_more_units_than_workers():
    """With more units than workers the queue keeps every worker busy: the makespan stays within one item of the
    work-conserving bound (what the static two-wave launch cannot do)."""
    rng = random.Random(3)
    n_units, n_segs, n_workers = 300, 4, 100
    t = simulate(n_units, n_segs, n_workers, rng)
    lower = n_units * n_segs * 1.0 / n_workers            # mean item duration is 1.0
    assert t <= lower * 1.08 + 1.5
