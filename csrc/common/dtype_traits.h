// C++ element type -> runtime descriptor used by the kernels and the bindings.
//
// Plays the role of mpi::get_datatype<T>() in the reference
// (aurora.mpich.miniapps/src/include/mpi_datatype.hpp:18-51: C++ type -> MPI_Datatype,
// default MPI_BYTE).  There is no MPI here; what a collective kernel needs to
// know about T is its size, whether an arithmetic reduction exists for it
// (float -> f32 adds / multimem.add.f32, int -> s32 adds) and a printable name.
// Types without a native reduction fall back to "bytes" (copy-only), the
// analogue of MPI_BYTE.
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
HPCP_DTYPE(unsigned int, "unsigned int", true, ElemType::kInt, "uint32");  // two's complement add
HPCP_DTYPE(double, "double", false, ElemType::kFloat, "float64");
HPCP_DTYPE(long, "long", false, ElemType::kInt, "int64");
HPCP_DTYPE(unsigned long, "unsigned long", false, ElemType::kInt, "uint64");
HPCP_DTYPE(short, "short", false, ElemType::kInt, "int16");
HPCP_DTYPE(unsigned short, "unsigned short", false, ElemType::kInt, "uint16");
HPCP_DTYPE(unsigned char, "unsigned char", false, ElemType::kInt, "uint8");
#undef HPCP_DTYPE

template <typename T>
constexpr DTypeInfo get_dtype(const T& = T{}) {
  return dtype_of<T>::value;
}

// Run-time lookup by name ("float" | "int"); returns false if unknown / not reducible.
inline bool elem_type_from_name(const std::string& name, ElemType* out) {
  if (name == "float" || name == "float32" || name == "f32") {
    *out = ElemType::kFloat;
    return true;
  }
  if (name == "int" || name == "int32" || name == "s32") {
    *out = ElemType::kInt;
    return true;
  }
  return false;
}
inline const char* elem_type_name(ElemType t) { return t == ElemType::kFloat ? "float" : "int"; }

}  // namespace hpcp
