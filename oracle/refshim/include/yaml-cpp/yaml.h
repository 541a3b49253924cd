// yaml-cpp/yaml.h — API SHIM, not yaml-cpp (test infrastructure, see Eigen/Dense in this tree).
// A Node is a tree built in code (oracle/refshim/ref_capi.cpp fills it with the keys of
// config/hilti_config.yaml); only operator[](key) and as<T>() exist.
#ifndef ESKF_REFSHIM_YAML_H_
#define ESKF_REFSHIM_YAML_H_

#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace YAML
{
class Node
{
public:
  Node() : d_(std::make_shared<Data>()) {}
  Node operator[](const std::string & key) const
  {
    auto it = d_->children.find(key);
    if (it == d_->children.end()) {throw std::runtime_error("yaml shim: missing key '" + key + "'");}
    return it->second;
  }
  Node & child(const std::string & key) {return d_->children[key];}  // creates
  void set(double v) {d_->values.assign(1, v);}
  void set(const std::vector<double> & v) {d_->values = v;}
  template<class T> T as() const {return Convert<T>::get(d_->values);}

private:
  struct Data
  {
    std::map<std::string, Node> children;
    std::vector<double> values;
  };
  template<class T> struct Convert
  {
    static T get(const std::vector<double> & v)
    {
      if (v.size() != 1) {throw std::runtime_error("yaml shim: not a scalar");}
      return static_cast<T>(v[0]);
    }
  };
  std::shared_ptr<Data> d_;
};
template<> struct Node::Convert<bool>
{
  static bool get(const std::vector<double> & v)
  {
    if (v.size() != 1) {throw std::runtime_error("yaml shim: not a scalar");}
    return v[0] != 0.0;
  }
};
template<> struct Node::Convert<std::vector<double>>
{
  static std::vector<double> get(const std::vector<double> & v) {return v;}
};
}  // namespace YAML

#endif  // ESKF_REFSHIM_YAML_H_
