// Oracle build shim: boost::split / is_any_of / lexical_cast as used by socket.cpp.
#ifndef COS_SHIM_BOOST_ALGORITHM_STRING_HPP_
#define COS_SHIM_BOOST_ALGORITHM_STRING_HPP_
#include <sstream>
#include <string>
#include <vector>
namespace boost {
struct any_of_pred { std::string set; };
inline any_of_pred is_any_of(const std::string& s) { any_of_pred p; p.set = s; return p; }
inline void split(std::vector<std::string>& out, const std::string& in, const any_of_pred& p) {
  out.clear();
  std::string cur;
  for (size_t i = 0; i < in.size(); ++i) {
    if (p.set.find(in[i]) != std::string::npos) { out.push_back(cur); cur.clear(); }
    else cur.push_back(in[i]);
  }
  out.push_back(cur);
}
template <typename T, typename S> T lexical_cast(const S& v) {
  std::stringstream ss; ss << v; T t; ss >> t; return t;
}
}  // namespace boost
#endif
