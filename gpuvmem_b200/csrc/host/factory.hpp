// factory.hpp — string/int-keyed self-registering factories.
// Same surface as the reference's include/factory.cuh:8-93 (Singleton<Factory<T,Id>>,
// registerCreationFunction<T,Id>(id, creator), createObject<T,Id>(id); an unknown id
// prints a message and exits with -1, no exception crosses the boundary).
#pragma once
#include <cstdlib>
#include <iostream>
#include <map>
#include <stdexcept>
#include <typeinfo>

namespace gpuvmem {

template <class T>
class Singleton {
 public:
  static T& Instance() {
    static T the_one;
    return the_one;
  }
  Singleton() = delete;
};

template <class AbstractProduct, class IdentifierType, class ProductCreator = AbstractProduct* (*)()>
class Factory {
 public:
  struct UnknownId : std::runtime_error {
    UnknownId() : std::runtime_error("Unknown object type passed to Factory") {}
  };
  bool Register(const IdentifierType& id, ProductCreator make) { return makers_.emplace(id, make).second; }
  bool Unregister(const IdentifierType& id) { return makers_.erase(id) == 1; }
  bool Has(const IdentifierType& id) const { return makers_.count(id) != 0; }
  AbstractProduct* CreateObject(const IdentifierType& id) const {
    auto hit = makers_.find(id);
    if (hit == makers_.end()) throw UnknownId();
    return hit->second();
  }

 private:
  std::map<IdentifierType, ProductCreator> makers_;
};

template <class T, class V>
T* createObject(V value) {
  try {
    return Singleton<Factory<T, V>>::Instance().CreateObject(value);
  } catch (std::exception& err) {
    std::cerr << err.what() << " of class " << typeid(T).name() << " and missing id: " << value << std::endl;
    std::exit(-1);
  }
}

template <class T, class V, class Creator = T* (*)()>
bool registerCreationFunction(V value, Creator function) {
  return Singleton<Factory<T, V>>::Instance().Register(value, function);
}

}  // namespace gpuvmem
