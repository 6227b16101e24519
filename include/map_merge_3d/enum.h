// enum.h — string-convertible enum classes with the same surface as the reference's
// ENUM_CLASS helper (map_merge_3d/include/map_merge_3d/enum.h:30-67): `enum class E`,
// enums::to_string(E), enums::from_string<E>(std::string) (exact-case, throws
// std::runtime_error on an unknown name) and operator<<.
#ifndef MM3D_SHIM_ENUM_H_
#define MM3D_SHIM_ENUM_H_

#include <ostream>
#include <stdexcept>
#include <string>
#include <type_traits>

namespace map_merge_3d
{
namespace enums
{
template <typename E>
struct Names;  // specialised per enum: static const char* const* list(); static int count();
}  // namespace enums
}  // namespace map_merge_3d

#define MM3D_ENUM_STR_(x) #x,
#define MM3D_ENUM_CLASS(EnumType, LIST)                                              \
  enum class EnumType { LIST(MM3D_ENUM_ID_) };                                       \
  namespace enums                                                                    \
  {                                                                                  \
  template <>                                                                        \
  struct Names<EnumType> {                                                           \
    static const char* const* list()                                                 \
    {                                                                                \
      static const char* const n[] = {LIST(MM3D_ENUM_STR_)};                         \
      return n;                                                                      \
    }                                                                                \
    static int count()                                                               \
    {                                                                                \
      static const char* const n[] = {LIST(MM3D_ENUM_STR_)};                         \
      return (int)(sizeof(n) / sizeof(n[0]));                                        \
    }                                                                                \
    static const char* type_name() { return #EnumType; }                             \
  };                                                                                 \
  inline const char* to_string(EnumType e)                                           \
  {                                                                                  \
    return Names<EnumType>::list()[static_cast<std::underlying_type_t<EnumType>>(e)]; \
  }                                                                                  \
  }                                                                                  \
  inline std::ostream& operator<<(std::ostream& s, EnumType v) { return s << enums::to_string(v); }
#define MM3D_ENUM_ID_(x) x,

namespace map_merge_3d
{
namespace enums
{
template <typename T>
std::enable_if_t<std::is_enum<T>::value, T> from_string(const std::string& s)
{
  for (int i = 0; i < Names<T>::count(); ++i)
    if (s == Names<T>::list()[i]) return static_cast<T>(i);
  // same text as the reference's ENUM_CLASS (enum.h:58-60)
  throw std::runtime_error("from_string: " + s + " is invalid value for enum " + Names<T>::type_name());
}
}  // namespace enums
}  // namespace map_merge_3d

#endif  // MM3D_SHIM_ENUM_H_
