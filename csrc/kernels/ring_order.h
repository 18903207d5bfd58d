// Slot and flow-control rules of the fused ring allreduce (ring_allreduce.cu) as plain functions, shared by
// the kernel and the host: tests/test_ring_protocol.py runs a randomly scheduled model of P ranks on top of
// exactly these rules and checks that no chunk is overwritten before its reader has consumed it.
//
// Hop t = 0 .. P-1 of a rank: read block `src` (t = 0: the rank's own VA, else receive slot ring_src_slot(t)),
// add it into VC and, unless it is the last hop, forward it into the right neighbour's slot ring_fwd_slot(t).
#pragma once

#if defined(__CUDACC__)
#define HPCP_RING_HD __host__ __device__ __forceinline__
#else
#define HPCP_RING_HD inline
#endif

namespace hpcp {

// two_slots = false: P-1 slots, hop t lands in slot t-1 (never reused inside one allreduce).
// two_slots = true : the reference's VA/VB double buffer, hop t lands in slot (t-1) % 2.
HPCP_RING_HD int ring_src_slot(int t, bool two_slots) { return two_slots ? ((t - 1) & 1) : t - 1; }
HPCP_RING_HD int ring_fwd_slot(int t, bool two_slots) { return two_slots ? (t & 1) : t; }
HPCP_RING_HD bool ring_forwards(int t, int world) { return t < world - 1; }
// Two slots: the forward of hop t >= 2 overwrites what the right neighbour reads at ITS hop t-1, so the sender
// first waits for the neighbour's ack of hop t-1 ...
HPCP_RING_HD bool ring_waits_for_ack(int t, int world) { return t >= 2 && ring_forwards(t, world); }
// ... and a receiver acks exactly the hops whose slot will be written again (by the sender's hop t+1).
HPCP_RING_HD bool ring_publishes_ack(int t, int world) { return t >= 1 && ring_waits_for_ack(t + 1, world); }

// ---- pull variant: a rank READS the block its left neighbour processed one hop earlier (NVLink loads), adds it and
// keeps a local copy for its right neighbour to read one hop later.  Hop 1 reads the neighbour's VA directly.
// Local copy of hop t (1 <= t <= P-2; the last hop is read by nobody): slot t-1, or (t-1) % 2 with two slots.
HPCP_RING_HD bool ring_pull_keeps_copy(int t, int world) { return t >= 1 && t <= world - 2; }
HPCP_RING_HD int ring_pull_copy_slot(int t, bool two_slots) { return two_slots ? ((t - 1) & 1) : t - 1; }
// Source of hop t >= 2 inside the LEFT neighbour's slots = the copy it kept at its hop t-1.
HPCP_RING_HD int ring_pull_src_slot(int t, bool two_slots) { return ring_pull_copy_slot(t - 1, two_slots); }
// Two slots: the copy of hop t overwrites the copy of hop t-2, which the right neighbour reads at ITS hop t-1 ...
HPCP_RING_HD bool ring_pull_waits_for_ack(int t, int world) { return t >= 3 && ring_pull_keeps_copy(t, world); }
// ... so a reader acks the hops whose source copy will be overwritten (by the owner's hop t+1).
HPCP_RING_HD bool ring_pull_publishes_ack(int t, int world) { return t >= 2 && ring_pull_waits_for_ack(t + 1, world); }

}  // namespace hpcp
