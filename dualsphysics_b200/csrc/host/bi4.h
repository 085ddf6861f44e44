// bi4.h -- DualSPHysics binary container (JBinaryData) reader/writer for Case.bi4 in, Part_XXXX.bi4 out.
// Format restated from src/source/JBinaryData.cpp:779-966,1095-1101,1283-1299 and the values/arrays
// JPartDataBi4 writes (src/source/JPartDataBi4.cpp:190-390); see also dualsphysics_b200/bi4.py.
#pragma once
#include <stdint.h>
#include <string>
#include <vector>
#include <memory>

namespace b2s {

enum Bi4Type { BT_TEXT=1,BT_BOOL=2,BT_CHAR=3,BT_UCHAR=4,BT_SHORT=5,BT_USHORT=6,BT_INT=7,BT_UINT=8,BT_LLONG=9,BT_ULLONG=10,
               BT_FLOAT=11,BT_DOUBLE=12,BT_INT3=20,BT_UINT3=21,BT_FLOAT3=22,BT_DOUBLE3=23 };

size_t bi4_type_size(int t);

struct Bi4Value {
  std::string name; int type;
  std::string text; int64_t i; uint64_t u; double d; double d3[3];   // storage by category
};
struct Bi4Array {
  std::string name; int type; uint64_t count;
  std::vector<char> data;        // owned (when loaded)
  const void* ext;               // external pointer (when saving)
  Bi4Array():type(0),count(0),ext(nullptr){}
  const void* ptr()const{ return ext? ext: (const void*)data.data(); }
};
struct Bi4Item {
  std::string name; bool hide, hidevalues; std::string fmtfloat, fmtdouble;
  std::vector<Bi4Value> values; std::vector<Bi4Array> arrays; std::vector<std::unique_ptr<Bi4Item>> items;
  explicit Bi4Item(const std::string& n=""):name(n),hide(false),hidevalues(false),fmtfloat("%.7E"),fmtdouble("%.15E"){}
  void set_text(const std::string& n,const std::string& v);
  void set_bool(const std::string& n,bool v);
  void set_int(const std::string& n,int v);
  void set_uint(const std::string& n,unsigned v);
  void set_ullong(const std::string& n,uint64_t v);
  void set_llong(const std::string& n,int64_t v);
  void set_float(const std::string& n,float v);
  void set_float3(const std::string& n,const float v[3]);
  void set_double(const std::string& n,double v);
  void set_double3(const std::string& n,const double v[3]);
  const Bi4Value* find(const std::string& n)const;
  double get_double(const std::string& n,double def=0)const;
  uint64_t get_uint(const std::string& n,uint64_t def=0)const;
  bool get_bool(const std::string& n,bool def=false)const;
  bool get_double3(const std::string& n,double out[3])const;
  const Bi4Array* array(const std::string& n)const;
  Bi4Item* item(const std::string& n)const;
  Bi4Item* add_item(const std::string& n);
  void add_array_ext(const std::string& n,int type,uint64_t count,const void* p);
};

// throw std::runtime_error on failure
std::unique_ptr<Bi4Item> bi4_load(const std::string& file);
void bi4_save(const std::string& file,const Bi4Item& root,const std::string& filecode);
// List files (JBinaryData::SaveFileListApp, JBinaryData.cpp:1736-1758): `head` is written when the file is created,
// `item` (may be NULL: head only) is appended behind whatever the file already holds.
void bi4_save_listapp(const std::string& file,const std::string& filecode,const Bi4Item& head,const Bi4Item* item,bool truncate);
// All top-level items of a list file: [0] = head, then the appended items.
std::vector<std::unique_ptr<Bi4Item>> bi4_load_list(const std::string& file);
void bi4_add_array_owned(Bi4Item& it,const std::string& n,int type,uint64_t count,const void* p);   // copies the data

} // namespace b2s
