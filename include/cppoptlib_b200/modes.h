// cppoptlib_b200/modes.h -- function_base.h:42-46, shared by cppoptlib.h and expressions.h.
#ifndef CPPOPTLIB_B200_MODES_H_
#define CPPOPTLIB_B200_MODES_H_
namespace cppoptlib {
namespace function {
enum class DifferentiabilityMode { None = 0, First = 1, Second = 2 };
}  // namespace function
}  // namespace cppoptlib
#endif  // CPPOPTLIB_B200_MODES_H_
