// DSQ container reader + host-side repacking into the device planes (see dsq.h).
#include "dsq.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstring>
#include <stdexcept>

namespace dsocr {

namespace {
struct Cursor {
  const uint8_t* p;
  const uint8_t* end;
  template <typename T> T get() {
    if (p + sizeof(T) > end) throw std::runtime_error("snapshot malformed: truncated header");
    T v; memcpy(&v, p, sizeof(T)); p += sizeof(T);
    return v;
  }
  std::string str() {
    const uint32_t n = get<uint32_t>();
    if (p + n > end) throw std::runtime_error("snapshot malformed: truncated string");
    std::string s((const char*)p, n); p += n;
    return s;
  }
};
DsqDType dtype_from(uint32_t v) {
  switch (v) {
    case 0: case 1: case 8: case 12: case 14: case 16: return static_cast<DsqDType>(v);
    default: throw std::runtime_error("snapshot malformed: unsupported tensor dtype code " + std::to_string(v));
  }
}
}  // namespace

DsqReader::DsqReader(const std::string& path) {
  fd_ = ::open(path.c_str(), O_RDONLY);
  if (fd_ < 0) throw std::runtime_error("cannot open snapshot " + path);
  struct stat st;
  if (fstat(fd_, &st) != 0) throw std::runtime_error("cannot stat snapshot " + path);
  size_ = (size_t)st.st_size;
  base_ = (const uint8_t*)mmap(nullptr, size_, PROT_READ, MAP_PRIVATE, fd_, 0);
  if (base_ == MAP_FAILED) { base_ = nullptr; throw std::runtime_error("mmap failed for snapshot " + path); }
  Cursor c{base_, base_ + size_};
  if (size_ < 11 || memcmp(base_, "DSQSNAP", 7) != 0) throw std::runtime_error("invalid snapshot magic");
  c.p += 7;
  const uint32_t version = c.get<uint32_t>();
  if (version != 1) throw std::runtime_error("unsupported snapshot version " + std::to_string(version) + ", expected 1");
  candle_version = c.str(); model_id = c.str(); backend = c.str();
  default_dtype_ = dtype_from(c.get<uint32_t>());
  const uint32_t block_size = c.get<uint32_t>();
  if (block_size == 0) throw std::runtime_error("snapshot validation failed: block_size must be non-zero");
  block_size_ = block_size;
  const uint32_t count = c.get<uint32_t>();
  records_.reserve(count);
  for (uint32_t i = 0; i < count; ++i) {
    DsqRecord r;
    r.name = c.str();
    r.out_dim = c.get<uint32_t>(); r.in_dim = c.get<uint32_t>();
    r.q_dtype = dtype_from(c.get<uint32_t>());
    r.q_offset = c.get<uint64_t>(); r.q_len = c.get<uint64_t>();
    r.bias_offset = c.get<uint64_t>(); r.bias_len = c.get<uint64_t>();
    r.bias_dtype = c.get<uint32_t>();
    r.has_bias = r.bias_len != 0;  // bias_len == 0 -> (None, None, None), lib.rs:352-357
    if (r.has_bias && r.bias_dtype > 6)  // DsqBiasDType::try_from: U8 U32 I64 F16 F32 F64 BF16 = 0..6
      throw std::runtime_error("snapshot malformed: unsupported bias dtype code " + std::to_string(r.bias_dtype));
    records_.push_back(std::move(r));
  }
  // validate_records + the duplicate check of DsqReader::from_mmap (crates/dsq/src/lib.rs:217-232, 409-520), same
  // order and wording.  Like the reference, `open` does not compare the payload length of a *block* dtype with its
  // dims (only in_dim % block); that is checked when the tensor is uploaded (load_quant).
  const uint64_t meta_len = (uint64_t)(c.p - base_);
  auto fail = [](const std::string& m) { throw std::runtime_error("snapshot validation failed: " + m); };
  // validate_header (lib.rs:393-407)
  if (dsq_block_elems(default_dtype_) == 0) fail("snapshot dtype " + std::to_string((uint32_t)default_dtype_) + " not supported");
  if ((int)block_size_ != dsq_block_elems(default_dtype_))
    fail("snapshot block size " + std::to_string(block_size_) + " mismatches expected " + std::to_string(dsq_block_elems(default_dtype_)));
  if (size_ < meta_len) fail("file smaller than metadata region");
  auto check_bounds = [&](uint64_t off, uint64_t len, const std::string& name, const char* label) {
    if (off + len < off) fail("tensor `" + name + "` " + label + " slice " + std::to_string(off) + "+" + std::to_string(len) + " overflows usize");
    if (off + len > size_)
      fail("tensor `" + name + "` " + label + " slice [" + std::to_string(off) + ", " + std::to_string(off) + "+" + std::to_string(len) +
           ") exceeds file size " + std::to_string(size_));
  };
  for (size_t i = 0; i < records_.size(); ++i) {
    const DsqRecord& r = records_[i];
    if (r.q_len == 0) fail("tensor `" + r.name + "` has empty quantized payload");
    if (r.q_offset < meta_len)
      fail("tensor `" + r.name + "` q_offset " + std::to_string(r.q_offset) + " overlaps metadata (" + std::to_string(meta_len) + " bytes)");
    check_bounds(r.q_offset, r.q_len, r.name, "quantized");
    if (r.has_bias) check_bounds(r.bias_offset, r.bias_len, r.name, "bias");
    const int be = dsq_block_elems(r.q_dtype);
    if (be) {
      if (r.in_dim % be)
        fail("tensor `" + r.name + "` in_dim " + std::to_string(r.in_dim) + " not divisible by block_size " + std::to_string(be));
    } else {
      const uint64_t es = r.q_dtype == DsqDType::F32 ? 4 : 2;
      const uint64_t expected = (uint64_t)r.out_dim * r.in_dim * es;
      if (r.q_len != expected)
        fail("tensor `" + r.name + "` has q_len " + std::to_string(r.q_len) + " but expected " + std::to_string(expected) + " bytes");
    }
  }
  for (size_t i = 0; i < records_.size(); ++i)
    if (!index_.emplace(records_[i].name, i).second) fail("duplicate tensor record `" + records_[i].name + "`");
}

DsqReader::~DsqReader() {
  if (base_) munmap((void*)base_, size_);
  if (fd_ >= 0) ::close(fd_);
}

const DsqRecord* DsqReader::find(const std::string& name) const {
  auto it = index_.find(name);
  return it == index_.end() ? nullptr : &records_[it->second];
}

void dsq_alloc(QuantWeight& w, DsqDType fmt, long long N, int K, int count) {
  w.fmt = fmt; w.N = N; w.K = K; w.count = count;
  const size_t rows = (size_t)N * count;
  switch (fmt) {
    case DsqDType::Q8_0: w.a.alloc(rows * K); w.b.alloc(rows * (K / 32) * 2); break;
    case DsqDType::Q4K: w.a.alloc(rows * (K / 256) * 144); break;
    case DsqDType::Q6K: w.a.alloc(rows * (K / 2)); w.b.alloc(rows * (K / 4)); w.c.alloc(rows * (K / 16)); w.d.alloc(rows * (K / 256) * 2); break;
    default: w.fmt = DsqDType::F32; w.a.alloc(rows * K * 4); break;
  }
}

void dsq_upload_rows(QuantWeight& dst, long long row0, const uint8_t* src, DsqDType src_fmt, long long rows) {
  const int K = dst.K;
  if ((dsq_block_elems(src_fmt) != 0) != (dst.fmt != DsqDType::F32) || (dsq_block_elems(src_fmt) && src_fmt != dst.fmt))
    throw std::runtime_error("dsq: mixed dtypes inside one stacked weight are not supported");
  if (src_fmt == DsqDType::Q8_0) {
    const size_t nb = (size_t)K / 32;
    std::vector<int8_t> qs((size_t)rows * K);
    std::vector<uint16_t> d((size_t)rows * nb);
    for (size_t b = 0; b < (size_t)rows * nb; ++b) {
      memcpy(&d[b], src + b * 34, 2);
      memcpy(&qs[b * 32], src + b * 34 + 2, 32);
    }
    h2d((uint8_t*)dst.a.p + (size_t)row0 * K, qs.data(), qs.size());
    h2d((uint8_t*)dst.b.p + (size_t)row0 * nb * 2, d.data(), d.size() * 2);
  } else if (src_fmt == DsqDType::Q4K) {
    const size_t bytes = (size_t)rows * (K / 256) * 144;
    h2d((uint8_t*)dst.a.p + (size_t)row0 * (K / 256) * 144, src, bytes);
  } else if (src_fmt == DsqDType::Q6K) {
    const size_t nb = (size_t)K / 256;
    std::vector<uint8_t> ql((size_t)rows * K / 2), qh((size_t)rows * K / 4), sc((size_t)rows * K / 16);
    std::vector<uint16_t> d((size_t)rows * nb);
    for (size_t b = 0; b < (size_t)rows * nb; ++b) {
      const uint8_t* blk = src + b * 210;
      memcpy(&ql[b * 128], blk, 128);
      memcpy(&qh[b * 64], blk + 128, 64);
      memcpy(&sc[b * 16], blk + 192, 16);
      memcpy(&d[b], blk + 208, 2);
    }
    h2d((uint8_t*)dst.a.p + (size_t)row0 * K / 2, ql.data(), ql.size());
    h2d((uint8_t*)dst.b.p + (size_t)row0 * K / 4, qh.data(), qh.size());
    h2d((uint8_t*)dst.c.p + (size_t)row0 * K / 16, sc.data(), sc.size());
    h2d((uint8_t*)dst.d.p + (size_t)row0 * nb * 2, d.data(), d.size() * 2);
  } else {
    const size_t n = (size_t)rows * K;
    std::vector<float> f(n);
    if (src_fmt == DsqDType::F32) memcpy(f.data(), src, n * 4);
    else {
      const uint16_t* p = (const uint16_t*)src;
      for (size_t i = 0; i < n; ++i) f[i] = f16_to_32(p[i], src_fmt == DsqDType::BF16 ? DType::BF16 : DType::F16);
    }
    h2d((uint8_t*)dst.a.p + (size_t)row0 * K * 4, f.data(), n * 4);
  }
}

}  // namespace dsocr
