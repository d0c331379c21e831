// Read-only safetensors container (mmap).  Header: u64 LE length + JSON {name: {dtype, shape, data_offsets}}.
#pragma once
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <map>
#include <string>
#include <vector>

#include "json.h"

namespace dsocr {

struct StTensor {
  std::string dtype;  // "F32" | "F16" | "BF16"
  std::vector<long long> shape;
  const uint8_t* data = nullptr;
  size_t bytes = 0;
  long long numel() const { long long n = 1; for (auto d : shape) n *= d; return n; }
};

class SafeTensors {
 public:
  explicit SafeTensors(const std::string& path) {
    try { open_and_index(path); } catch (...) { release(); throw; }  // the destructor does not run when the constructor throws
  }
  ~SafeTensors() { release(); }
  SafeTensors(const SafeTensors&) = delete;
  bool has(const std::string& n) const { return tensors_.count(n) != 0; }
  const StTensor& get(const std::string& n) const {
    auto it = tensors_.find(n);
    if (it == tensors_.end()) throw std::runtime_error("missing tensor `" + n + "` in checkpoint");
    return it->second;
  }

 private:
  void release() {
    if (base_ && base_ != MAP_FAILED) munmap((void*)base_, size_);
    if (fd_ >= 0) ::close(fd_);
    base_ = nullptr; fd_ = -1;
  }
  void open_and_index(const std::string& path) {
    fd_ = ::open(path.c_str(), O_RDONLY);
    if (fd_ < 0) throw std::runtime_error("cannot open weights file " + path);
    struct stat st;
    if (fstat(fd_, &st) != 0) throw std::runtime_error("cannot stat " + path);
    size_ = (size_t)st.st_size;
    if (size_ < 8) throw std::runtime_error("safetensors file too small: " + path);
    base_ = (const uint8_t*)mmap(nullptr, size_, PROT_READ, MAP_PRIVATE, fd_, 0);
    if (base_ == MAP_FAILED) throw std::runtime_error("mmap failed for " + path);
    uint64_t hlen = 0;
    memcpy(&hlen, base_, 8);
    if (hlen > size_ - 8) throw std::runtime_error("safetensors header length exceeds file size");  // no 8 + hlen wrap-around
    // the parser may scan a number past the end of its range (strtod): hand it a NUL-terminated copy of the header
    const std::string header((const char*)base_ + 8, (size_t)hlen);
    Json hdr = JsonParser(header.c_str(), header.size()).parse();
    const size_t payload = size_ - 8 - (size_t)hlen;
    const uint8_t* data0 = base_ + 8 + hlen;
    for (auto& kv : hdr.obj) {
      if (kv.first == "__metadata__") continue;
      StTensor t;
      t.dtype = kv.second.at("dtype").str;
      for (auto& d : kv.second.at("shape").arr) t.shape.push_back((long long)d.num);
      const auto& off = kv.second.at("data_offsets").arr;
      const size_t a = (size_t)off.at(0).num, b = (size_t)off.at(1).num;
      if (b < a || b > payload) throw std::runtime_error("safetensors: bad offsets for " + kv.first);
      t.data = data0 + a;
      t.bytes = b - a;
      const size_t el = t.dtype == "F32" ? 4 : (t.dtype == "F16" || t.dtype == "BF16") ? 2 : 0;
      long long n = 1;
      for (auto d : t.shape) { if (d < 0 || (d > 0 && n > (long long)(payload / (size_t)d) + 1)) throw std::runtime_error("safetensors: bad shape for " + kv.first); n *= d; }
      // readers take numel * element size bytes from the mapping: the header must agree with itself
      if (el && (size_t)n * el != t.bytes) throw std::runtime_error("safetensors: byte length of " + kv.first + " does not match its shape");
      tensors_.emplace(kv.first, std::move(t));
    }
  }
  int fd_ = -1;
  size_t size_ = 0;
  const uint8_t* base_ = nullptr;
  std::map<std::string, StTensor> tensors_;
};

}  // namespace dsocr
