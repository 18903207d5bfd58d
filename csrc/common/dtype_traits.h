// C++ element type -> runtime descriptor used by the kernels and the bindings.
//
// Plays the role of mpi::get_datatype<T>() in the reference
// (aurora.mpich.miniapps/src/include/mpi_datatype.hpp:18-51: C++ type -> MPI_Datatype,
// default MPI_BYTE).  There is no MPI here; what a collective kernel needs to
// know about T is its size, whether an arithmetic reduction exists for it and a
// printable name.  Every type the reference's trait maps to an MPI_SUM-capable
// datatype (short/int/long/float/double and the unsigned integers, :28-51) has
// SUM kernels here (ring, accumulate, two-shot: csrc/kernels/ring_allreduce.cu);
// `long double` has no device representation on NVIDIA GPUs and, like any other
// type, falls back to "bytes" (copy-only), the analogue of MPI_BYTE.
#pragma once

#include <cstddef>
#include <cstdint>
#include <string>

#include "../kernels/api.h"

namespace hpcp {

struct DTypeInfo {
  const char* name;      // "float", "int", ... ; "bytes" for opaque types
  size_t size;           // sizeof(T)
  bool reducible;        // has a SUM kernel in this suite
  ElemType elem;         // valid iff reducible
  const char* torch;     // matching torch dtype name, "" if none
};

template <typename T>
struct dtype_of {
  static constexpr DTypeInfo value{"bytes", sizeof(T), false, ElemType::kFloat, ""};
};

#define HPCP_DTYPE(T, NAME, RED, ELEM, TORCH)                                 \
  template <>                                                                 \
  struct dtype_of<T> {                                                        \
    static constexpr DTypeInfo value{NAME, sizeof(T), RED, ELEM, TORCH};      \
  }

HPCP_DTYPE(float, "float", true, ElemType::kFloat, "float32");
HPCP_DTYPE(int, "int", true, ElemType::kInt, "int32");
HPCP_DTYPE(unsigned int, "unsigned int", true, ElemType::kUInt, "uint32");
HPCP_DTYPE(double, "double", true, ElemType::kDouble, "float64");
HPCP_DTYPE(long, "long", true, ElemType::kLong, "int64");
HPCP_DTYPE(long long, "long", true, ElemType::kLong, "int64");
HPCP_DTYPE(unsigned long, "unsigned long", true, ElemType::kULong, "uint64");
HPCP_DTYPE(unsigned long long, "unsigned long", true, ElemType::kULong, "uint64");
HPCP_DTYPE(short, "short", true, ElemType::kShort, "int16");
HPCP_DTYPE(unsigned short, "unsigned short", true, ElemType::kUShort, "uint16");
HPCP_DTYPE(unsigned char, "unsigned char", true, ElemType::kUChar, "uint8");
#undef HPCP_DTYPE

template <typename T>
constexpr DTypeInfo get_dtype(const T& = T{}) {
  return dtype_of<T>::value;
}

// Run-time lookup by name (C++ spelling, torch spelling or a short alias); false if unknown.
inline bool elem_type_from_name(const std::string& name, ElemType* out) {
  struct Row {
    const char* name;
    ElemType type;
  };
  static const Row rows[] = {
      {"float", ElemType::kFloat},    {"float32", ElemType::kFloat},       {"f32", ElemType::kFloat},
      {"int", ElemType::kInt},        {"int32", ElemType::kInt},           {"s32", ElemType::kInt},
      {"uint", ElemType::kUInt},      {"unsigned", ElemType::kUInt},       {"unsigned int", ElemType::kUInt},
      {"uint32", ElemType::kUInt},    {"u32", ElemType::kUInt},
      {"double", ElemType::kDouble},  {"float64", ElemType::kDouble},      {"f64", ElemType::kDouble},
      {"long", ElemType::kLong},      {"int64", ElemType::kLong},          {"s64", ElemType::kLong},
      {"ulong", ElemType::kULong},    {"unsigned long", ElemType::kULong}, {"uint64", ElemType::kULong},
      {"u64", ElemType::kULong},
      {"short", ElemType::kShort},    {"int16", ElemType::kShort},         {"s16", ElemType::kShort},
      {"ushort", ElemType::kUShort},  {"unsigned short", ElemType::kUShort}, {"uint16", ElemType::kUShort},
      {"u16", ElemType::kUShort},
      {"uchar", ElemType::kUChar},    {"unsigned char", ElemType::kUChar}, {"uint8", ElemType::kUChar},
      {"u8", ElemType::kUChar},
  };
  for (const Row& r : rows)
    if (name == r.name) {
      *out = r.type;
      return true;
    }
  return false;
}
inline const char* elem_type_name(ElemType t) {
  switch (t) {
    case ElemType::kFloat: return "float";
    case ElemType::kInt: return "int";
    case ElemType::kUInt: return "uint";
    case ElemType::kDouble: return "double";
    case ElemType::kLong: return "long";
    case ElemType::kULong: return "ulong";
    case ElemType::kShort: return "short";
    case ElemType::kUShort: return "ushort";
    case ElemType::kUChar: return "uchar";
  }
  return "?";
}

}  // namespace hpcp
