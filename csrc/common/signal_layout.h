// Signal-pad layout and status codes: the part of the cross-GPU protocol that host code, device code
// and the Python bindings all need (the device-side operations are in signal.cuh).
//
// Every rank owns a pad of 32-bit words in peer-mapped memory; words are monotonic epochs
// (see signal.cuh).  The fixed part is followed by `extra_words` per-chunk arrival words and then
// the rank's status word (NodeMemory::alloc_pads, parallel/symmetric.py::SignalPads).
#pragma once

#include <stdint.h>

namespace hpcp {

// Status codes written to the per-rank device status word.
enum : uint32_t {
  kStatusOk = 0,
  kStatusTimeout = 0x7100DEAD,   // a spin-wait hit its deadline
  kStatusMismatch = 0x0BADDA7A,  // fused verification found wrong payload
};

// Fixed pad layout (in 32-bit words).  kMaxRanks peers per section.
constexpr int kMaxRanks = 16;
constexpr int kPadBarrier = 0;                  // [0,16)   barrier arrival words
constexpr int kPadReady = kPadBarrier + 16;     // [16,32)  "receive posted" words (rendezvous)
constexpr int kPadDone = kPadReady + 16;        // [32,48)  "data landed" words
constexpr int kPadAck = kPadDone + 16;          // [48,64)  "data consumed" words
constexpr int kPadLocal = kPadAck + 16;         // [64,128) rank-local counters (CTA tickets)
constexpr int kPadWords = 128;                  // fixed part; chunk flags follow
constexpr int kPadChunkBase = kPadWords;        // per-chunk arrival words start here
constexpr int kPadTailWords = 32;               // room after the chunk words; its first word is the status word

}  // namespace hpcp
