// Test infrastructure (not product code): writes the text archive of a gpmp2::SignedDistanceField with a REAL
// Boost.Serialization runtime doing all the archive framing -- header, class-info emission (tracking level + version, once
// per class), token delimiters, item_version -- so that gpmp2_b200/boost_archive.py's restatement of that grammar is pinned
// to something other than memory.  This image has no Boost headers, only Nsight Compute's bundled
// libboost_serialization.so.1.78.0; the declarations below restate just enough of Boost 1.78's class interfaces
// (boost/archive/detail/basic_oserializer.hpp, basic_oarchive.hpp, boost/serialization/extended_type_info.hpp,
// boost/archive/basic_archive.hpp) to link against its exported symbols with the same object layouts and vtable order
// (checked against the library's vtable sizes: 7 virtual slots each).  What the runtime decides: everything
// basic_oarchive::save_object, end_preamble, newtoken, init and save(item_version_type) emit.  What is still restated here:
// the members each class serializes (SignedDistanceField.h:201-208; gtsam 4.0 Point3 -> base Vector3; gtsam/base/Matrix.h
// rows, cols, coefficient array) and the inline primitive formatting of basic_text_oprimitive (os << t; doubles with
// setprecision(max_digits10) << scientific).
//
// The binary archive likewise: binary_oarchive_impl's constructor / init (the 40-byte header), the class-info bytes
// (tracking_type 1 byte, version_type 4 bytes) come from the runtime; size_t / double members, collection_size_type
// (size_t) and item_version_type (unsigned int) go through its exported save_binary with the sizes those wrapper types
// have in Boost 1.78's headers, and the coefficient array as one block (array optimisation of binary archives).
//
//   boost_sdf_probe read text|bin <file>  ->  the runtime's basic_iarchive reads an archive (header check, class infos via
//   load_object) and the field it found is printed: proves that files written by saveSDF (stamped library version 17) are
//   accepted by a real Boost reader
//   boost_sdf_probe xmlname <tag>  ->  "accepted" or "threw: <what>" from the runtime's basic_xml_oarchive::save_start
//   boost_sdf_probe text|bin <rows> <cols> <nz>  ->  stdout: the archive of a field with origin (-1.5, 0.25, 2), cell 0.5
//   and data[z](r, c) = (100 z + 10 r + c) / 7
#include <cstdarg>
#include <exception>
#include <fstream>
#include <iterator>
#include <stdexcept>
#include <iomanip>
#include <iostream>
#include <limits>
#include <new>
#include <sstream>
#include <string>
#include <cstdlib>
#include <vector>

namespace boost {
namespace serialization {
class extended_type_info {
  const unsigned int m_type_info_key;
  virtual bool is_less_than(const extended_type_info&) const = 0;
  virtual bool is_equal(const extended_type_info&) const = 0;
  const char* m_key;
 protected:
  extended_type_info(const unsigned int type_info_key, const char* key);   // exported
  virtual ~extended_type_info();                                            // exported
 public:
  virtual const char* get_debug_info() const = 0;
  bool operator<(const extended_type_info& rhs) const;                      // exported
  virtual void* construct(unsigned int = 0, ...) const = 0;
  virtual void destroy(void const* const) const = 0;
};
struct item_version_type { unsigned int t; };
}  // namespace serialization
namespace archive {
class version_type {   // user-provided copy constructor, as in basic_archive.hpp: returned through memory
  unsigned int t;
 public:
  version_type() : t(0) {}
  explicit version_type(const unsigned int& t_) : t(t_) {}
  version_type(const version_type& o) : t(o.t) {}
};
namespace detail {
class basic_oarchive;
class basic_pointer_oserializer;
class basic_serializer {
  const boost::serialization::extended_type_info* m_eti;
 protected:
  explicit basic_serializer(const boost::serialization::extended_type_info& eti) : m_eti(&eti) {}
};
class basic_oserializer : public basic_serializer {
  basic_pointer_oserializer* m_bpos;
 protected:
  explicit basic_oserializer(const boost::serialization::extended_type_info& type);   // exported
  virtual ~basic_oserializer();                                                        // exported
 public:
  virtual void save_object_data(basic_oarchive& ar, const void* x) const = 0;
  virtual bool class_info() const = 0;
  virtual bool tracking(const unsigned int flags) const = 0;
  virtual version_type version() const = 0;
  virtual bool is_polymorphic() const = 0;
};
class basic_oarchive {
 public:
  void save_object(const void* x, const basic_oserializer& bos);   // exported
  void end_preamble();                                              // exported
};
class basic_iarchive;
class basic_pointer_iserializer;
class basic_iserializer : public basic_serializer {
  basic_pointer_iserializer* m_bpis;
 protected:
  explicit basic_iserializer(const boost::serialization::extended_type_info& type);   // exported
  virtual ~basic_iserializer();                                                        // exported
 public:
  virtual void load_object_data(basic_iarchive& ar, void* x, const unsigned int file_version) const = 0;
  virtual bool class_info() const = 0;
  virtual bool tracking(const unsigned int) const = 0;
  virtual version_type version() const = 0;
  virtual bool is_polymorphic() const = 0;
  virtual void destroy(void* address) const = 0;
};
class basic_iarchive {
 public:
  void load_object(void* t, const basic_iserializer& bis);         // exported
};
}  // namespace detail
class text_oarchive;
template <class A> class basic_text_oarchive { public: void newtoken(); void init(); };
template <class A> class text_oarchive_impl {
 public:
  text_oarchive_impl(std::ostream&, unsigned int);
  ~text_oarchive_impl();
  void save(const boost::serialization::item_version_type&);
};
class text_iarchive;
template <class A> class text_iarchive_impl {
 public:
  text_iarchive_impl(std::istream&, unsigned int);
  void init();                                                       // reads and checks the header
  void load(boost::serialization::item_version_type&);
};
class binary_iarchive;
template <class A, class E, class T> class basic_binary_iprimitive { public: void load_binary(void*, std::size_t); };
template <class A, class E, class T> class binary_iarchive_impl {
 public:
  binary_iarchive_impl(std::istream&, unsigned int);
  void init(unsigned int);
};
class xml_oarchive;
template <class A> class basic_xml_oarchive { public: void save_start(const char*); };
template <class A> class xml_oarchive_impl { public: xml_oarchive_impl(std::ostream&, unsigned int); };
class binary_oarchive;
template <class A, class E, class T> class basic_binary_oprimitive { public: void save_binary(const void*, std::size_t); };
template <class A, class E, class T> class binary_oarchive_impl {
 public:
  binary_oarchive_impl(std::ostream&, unsigned int);
  void init(unsigned int);
};
}  // namespace archive
}  // namespace boost

using boost::archive::detail::basic_oarchive;
typedef boost::archive::text_oarchive_impl<boost::archive::text_oarchive> TextImpl;
typedef boost::archive::basic_text_oarchive<boost::archive::text_oarchive> TextBase;
typedef boost::archive::binary_oarchive_impl<boost::archive::binary_oarchive, char, std::char_traits<char> > BinImpl;
typedef boost::archive::basic_binary_oprimitive<boost::archive::binary_oarchive, char, std::char_traits<char> > BinPrim;
static bool g_bin = false;
// binary_oarchive_impl : basic_binary_oprimitive, basic_binary_oarchive -- the polymorphic base is primary (offset 0), the
// primitive follows basic_oarchive's 0x28 bytes (the library's own code reads its streambuf reference at 0x28(this))
static BinPrim& bin_prim(basic_oarchive& ar) { return *reinterpret_cast<BinPrim*>(reinterpret_cast<char*>(&ar) + 0x28); }

static std::ostringstream g_os;

// one extended_type_info per serialized class (ordered by address, which is all basic_oarchive's class table needs)
struct Eti : boost::serialization::extended_type_info {
  Eti() : extended_type_info(77, nullptr) {}
  bool is_less_than(const extended_type_info& rhs) const override { return this < &rhs; }
  bool is_equal(const extended_type_info& rhs) const override { return this == &rhs; }
  const char* get_debug_info() const override { return "probe"; }
  void* construct(unsigned int, ...) const override { return nullptr; }
  void destroy(void const* const) const override {}
};

// primitives the way common_oarchive / text_oarchive_impl save them: end_preamble(), newtoken(), os << t
template <class T> static void prim(basic_oarchive& ar, const T& t) {
  ar.end_preamble();
  if (g_bin) { bin_prim(ar).save_binary(&t, sizeof(T)); return; }
  reinterpret_cast<TextBase&>(ar).newtoken();
  g_os << t;
}
static void prim(basic_oarchive& ar, double t) {
  ar.end_preamble();
  if (g_bin) { bin_prim(ar).save_binary(&t, sizeof(double)); return; }
  reinterpret_cast<TextBase&>(ar).newtoken();
  g_os << std::setprecision(std::numeric_limits<double>::max_digits10) << std::scientific << t;
}

// object_serializable, track_never... the defaults of a class without BOOST_CLASS_* macros: class info written,
// tracked only if serialized through a pointer (never here), version 0
struct Ser : boost::archive::detail::basic_oserializer {
  void (*fn)(basic_oarchive&, const void*);
  Ser(const Eti& e, void (*f)(basic_oarchive&, const void*)) : basic_oserializer(e), fn(f) {}
  void save_object_data(basic_oarchive& ar, const void* x) const override { fn(ar, x); }
  bool class_info() const override { return true; }
  bool tracking(const unsigned int) const override { return false; }
  boost::archive::version_type version() const override { return boost::archive::version_type(0); }
  bool is_polymorphic() const override { return false; }
};

struct Mat { size_t rows, cols; std::vector<double> a; };   // column-major, like Eigen
struct Sdf { double origin[3]; size_t rows, cols, nz; double cell; std::vector<Mat> data; };

static void save_eigen(basic_oarchive& ar, size_t rows, size_t cols, const double* a) {   // gtsam/base/Matrix.h save()
  prim(ar, rows);
  prim(ar, cols);
  if (g_bin) { ar.end_preamble(); bin_prim(ar).save_binary(a, rows * cols * sizeof(double)); return; }   // save_array: one block
  for (size_t i = 0; i < rows * cols; i++) prim(ar, a[i]);                               // make_array, element by element in text
}
static Eti e_sdf, e_pt, e_v3, e_vec, e_mat;
static void ser_v3(basic_oarchive& ar, const void* x) { save_eigen(ar, 3, 1, static_cast<const double*>(x)); }
static Ser s_v3(e_v3, ser_v3);
static void ser_pt(basic_oarchive& ar, const void* x) { ar.save_object(x, s_v3); }      // Point3: base_object<Vector3>
static Ser s_pt(e_pt, ser_pt);
static void ser_mat(basic_oarchive& ar, const void* x) {
  const Mat& m = *static_cast<const Mat*>(x);
  save_eigen(ar, m.rows, m.cols, m.a.data());
}
static Ser s_mat(e_mat, ser_mat);
static void ser_vec(basic_oarchive& ar, const void* x) {                                  // boost/serialization/vector.hpp
  const std::vector<Mat>& v = *static_cast<const std::vector<Mat>*>(x);
  prim(ar, v.size());                                                                     // collection_size_type
  boost::serialization::item_version_type iv{0};
  if (g_bin) { prim(ar, iv.t); }                                                          // unsigned int wrapper: 4 bytes
  else { ar.end_preamble(); reinterpret_cast<TextImpl&>(ar).save(iv); }                   // exported: newtoken + value
  for (const Mat& m : v) ar.save_object(&m, s_mat);
}
static Ser s_vec(e_vec, ser_vec);
static void ser_sdf(basic_oarchive& ar, const void* x) {                                  // SignedDistanceField.h:201-208
  const Sdf& s = *static_cast<const Sdf*>(x);
  ar.save_object(s.origin, s_pt);
  prim(ar, s.rows);
  prim(ar, s.cols);
  prim(ar, s.nz);
  prim(ar, s.cell);
  ar.save_object(&s.data, s_vec);
}
static Ser s_sdf(e_sdf, ser_sdf);

// ---- reading: the runtime's basic_iarchive::load_object consumes the class infos, the members are read as written ----
using boost::archive::detail::basic_iarchive;
typedef boost::archive::text_iarchive_impl<boost::archive::text_iarchive> TextIn;
typedef boost::archive::binary_iarchive_impl<boost::archive::binary_iarchive, char, std::char_traits<char> > BinIn;
typedef boost::archive::basic_binary_iprimitive<boost::archive::binary_iarchive, char, std::char_traits<char> > BinInPrim;
static std::istringstream g_is;
static BinInPrim& bin_in(basic_iarchive& ar) { return *reinterpret_cast<BinInPrim*>(reinterpret_cast<char*>(&ar) + 0x28); }
template <class T> static void rd(basic_iarchive& ar, T& t) {
  if (g_bin) { bin_in(ar).load_binary(&t, sizeof(T)); return; }
  if (!(g_is >> t)) throw std::runtime_error("input stream error");   // basic_text_iprimitive::load
}
struct ISer : boost::archive::detail::basic_iserializer {
  void (*fn)(basic_iarchive&, void*);
  ISer(const Eti& e, void (*f)(basic_iarchive&, void*)) : basic_iserializer(e), fn(f) {}
  void load_object_data(basic_iarchive& ar, void* x, const unsigned int) const override { fn(ar, x); }
  bool class_info() const override { return true; }
  bool tracking(const unsigned int) const override { return false; }
  boost::archive::version_type version() const override { return boost::archive::version_type(0); }
  bool is_polymorphic() const override { return false; }
  void destroy(void*) const override {}
};
static void load_eigen(basic_iarchive& ar, size_t& rows, size_t& cols, std::vector<double>& a) {
  rd(ar, rows);
  rd(ar, cols);
  if (rows * cols > (1u << 24)) throw std::runtime_error("implausible matrix size");
  a.resize(rows * cols);
  if (g_bin) { bin_in(ar).load_binary(a.data(), a.size() * sizeof(double)); return; }
  for (double& v : a) rd(ar, v);
}
static Eti ie_sdf, ie_pt, ie_v3, ie_vec, ie_mat;
static void ild_v3(basic_iarchive& ar, void* x) {
  size_t r, c;
  std::vector<double> a;
  load_eigen(ar, r, c, a);
  if (r != 3 || c != 1) throw std::runtime_error("Vector3 shape");
  for (int i = 0; i < 3; i++) static_cast<double*>(x)[i] = a[i];
}
static ISer is_v3(ie_v3, ild_v3);
static void ild_pt(basic_iarchive& ar, void* x) { ar.load_object(x, is_v3); }
static ISer is_pt(ie_pt, ild_pt);
static void ild_mat(basic_iarchive& ar, void* x) {
  Mat& m = *static_cast<Mat*>(x);
  load_eigen(ar, m.rows, m.cols, m.a);
}
static ISer is_mat(ie_mat, ild_mat);
static void ild_vec(basic_iarchive& ar, void* x) {
  std::vector<Mat>& v = *static_cast<std::vector<Mat>*>(x);
  size_t count;
  rd(ar, count);
  boost::serialization::item_version_type iv{0};
  if (g_bin) rd(ar, iv.t); else reinterpret_cast<TextIn&>(ar).load(iv);     // library version > 3: item_version present
  if (count > (1u << 20)) throw std::runtime_error("implausible layer count");
  v.resize(count);
  for (Mat& m : v) ar.load_object(&m, is_mat);
}
static ISer is_vec(ie_vec, ild_vec);
static void ild_sdf(basic_iarchive& ar, void* x) {
  Sdf& s = *static_cast<Sdf*>(x);
  ar.load_object(s.origin, is_pt);
  rd(ar, s.rows);
  rd(ar, s.cols);
  rd(ar, s.nz);
  rd(ar, s.cell);
  ar.load_object(&s.data, is_vec);
}
static ISer is_sdf(ie_sdf, ild_sdf);

alignas(64) static char g_archive[65536];   // the library's text_oarchive object lives here (size unknown without headers)

int main(int argc, char** argv) {
  if (argc == 3 && std::string(argv[1]) == "xmlname") {
    // does the runtime's xml_oarchive accept this tag name?  (the reference's .xml branch uses BOOST_SERIALIZATION_NVP(*this))
    new (g_archive) boost::archive::xml_oarchive_impl<boost::archive::xml_oarchive>(g_os, 0);
    try {
      reinterpret_cast<boost::archive::basic_xml_oarchive<boost::archive::xml_oarchive>*>(g_archive)->save_start(argv[2]);
      std::cout << "accepted\n";
    } catch (const std::exception& e) {
      std::cout << "threw: " << e.what() << "\n";
    }
    return 0;
  }
  if (argc == 4 && std::string(argv[1]) == "read") {
    // the runtime reads an archive (ours or its own): header check, class infos, then prints what it found
    g_bin = std::string(argv[2]) == "bin";
    std::ifstream f(argv[3], std::ios::binary);
    if (!f.good()) return 3;
    g_is.str(std::string((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>()));
    Sdf s{};
    try {
      if (g_bin) {
        new (g_archive) BinIn(g_is, 1);                                  // no_header: init explicitly, once
        reinterpret_cast<BinIn*>(g_archive)->init(0);
      } else {
        new (g_archive) TextIn(g_is, 0);
        reinterpret_cast<TextIn*>(g_archive)->init();
      }
      reinterpret_cast<basic_iarchive*>(g_archive)->load_object(&s, is_sdf);
    } catch (const std::exception& e) {
      std::cout << "threw: " << e.what() << "\n";
      return 1;
    }
    std::cout << std::setprecision(17) << s.rows << ' ' << s.cols << ' ' << s.nz << ' ' << s.cell << ' ' << s.origin[0] << ' '
              << s.origin[1] << ' ' << s.origin[2] << "\n";
    for (const Mat& m : s.data) {
      if (m.rows != s.rows || m.cols != s.cols) { std::cout << "layer shape mismatch\n"; return 1; }
      for (double v : m.a) std::cout << v << ' ';
      std::cout << "\n";
    }
    return 0;
  }
  if (argc != 5) return 2;
  g_bin = std::string(argv[1]) == "bin";
  Sdf s{{-1.5, 0.25, 2.0}, (size_t)atoi(argv[2]), (size_t)atoi(argv[3]), (size_t)atoi(argv[4]), 0.5, {}};
  for (size_t z = 0; z < s.nz; z++) {
    Mat m{s.rows, s.cols, std::vector<double>(s.rows * s.cols)};
    for (size_t r = 0; r < s.rows; r++)
      for (size_t c = 0; c < s.cols; c++) m.a[c * s.rows + r] = (100.0 * z + 10.0 * r + c) / 7.0;
    s.data.push_back(m);
  }
  if (g_bin) {
    new (g_archive) BinImpl(g_os, 0);                                   // writes the header itself (init(flags) inside)
    if (g_os.str().empty()) reinterpret_cast<BinImpl*>(g_archive)->init(0);
    reinterpret_cast<basic_oarchive*>(g_archive)->save_object(&s, s_sdf);
    g_os.flush();
    std::cout << g_os.str();                                            // (object deliberately not destroyed: no exported destructor)
    return 0;
  }
  TextImpl* impl = new (g_archive) TextImpl(g_os, 0);
  // Itanium ABI: the polymorphic base (basic_text_oarchive -> basic_oarchive) is the primary base at offset 0
  reinterpret_cast<TextBase*>(g_archive)->init();                      // text_oarchive's constructor does this: header
  reinterpret_cast<basic_oarchive*>(g_archive)->save_object(&s, s_sdf);   // oa << *this
  impl->~TextImpl();
  std::cout << g_os.str();
  return 0;
}
