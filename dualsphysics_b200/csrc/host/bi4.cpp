// bi4.cpp -- see bi4.h.
#include "bi4.h"
#include <stdexcept>
#include <cstring>
#include <cstdio>
#include <fstream>

namespace b2s {

size_t bi4_type_size(int t){
  switch(t){
    case BT_CHAR: case BT_UCHAR: return 1;
    case BT_SHORT: case BT_USHORT: return 2;
    case BT_BOOL: case BT_INT: case BT_UINT: case BT_FLOAT: return 4;
    case BT_LLONG: case BT_ULLONG: case BT_DOUBLE: return 8;
    case BT_INT3: case BT_UINT3: case BT_FLOAT3: return 12;
    case BT_DOUBLE3: return 24;
    default: return 0;
  }
}

static Bi4Value* upsert(std::vector<Bi4Value>& vs,const std::string& n){
  for(auto& v:vs)if(v.name==n)return &v;
  vs.emplace_back(); Bi4Value& v=vs.back(); v.name=n; v.type=0; v.i=0; v.u=0; v.d=0; v.d3[0]=v.d3[1]=v.d3[2]=0; return &v;
}
void Bi4Item::set_text(const std::string& n,const std::string& v){ Bi4Value* x=upsert(values,n); x->type=BT_TEXT; x->text=v; }
void Bi4Item::set_bool(const std::string& n,bool v){ Bi4Value* x=upsert(values,n); x->type=BT_BOOL; x->i=v?1:0; }
void Bi4Item::set_int(const std::string& n,int v){ Bi4Value* x=upsert(values,n); x->type=BT_INT; x->i=v; }
void Bi4Item::set_uint(const std::string& n,unsigned v){ Bi4Value* x=upsert(values,n); x->type=BT_UINT; x->u=v; }
void Bi4Item::set_ullong(const std::string& n,uint64_t v){ Bi4Value* x=upsert(values,n); x->type=BT_ULLONG; x->u=v; }
void Bi4Item::set_llong(const std::string& n,int64_t v){ Bi4Value* x=upsert(values,n); x->type=BT_LLONG; x->i=v; }
void Bi4Item::set_float(const std::string& n,float v){ Bi4Value* x=upsert(values,n); x->type=BT_FLOAT; x->d=v; }
void Bi4Item::set_float3(const std::string& n,const float v[3]){ Bi4Value* x=upsert(values,n); x->type=BT_FLOAT3; x->d3[0]=v[0]; x->d3[1]=v[1]; x->d3[2]=v[2]; }
void Bi4Item::set_double(const std::string& n,double v){ Bi4Value* x=upsert(values,n); x->type=BT_DOUBLE; x->d=v; }
void Bi4Item::set_double3(const std::string& n,const double v[3]){ Bi4Value* x=upsert(values,n); x->type=BT_DOUBLE3; x->d3[0]=v[0]; x->d3[1]=v[1]; x->d3[2]=v[2]; }
const Bi4Value* Bi4Item::find(const std::string& n)const{ for(auto& v:values)if(v.name==n)return &v; return nullptr; }
double Bi4Item::get_double(const std::string& n,double def)const{
  const Bi4Value* v=find(n); if(!v)return def;
  switch(v->type){ case BT_DOUBLE: case BT_FLOAT: return v->d; case BT_UINT: case BT_ULLONG: case BT_USHORT: case BT_UCHAR: return double(v->u); default: return double(v->i); }
}
uint64_t Bi4Item::get_uint(const std::string& n,uint64_t def)const{
  const Bi4Value* v=find(n); if(!v)return def;
  switch(v->type){ case BT_UINT: case BT_ULLONG: case BT_USHORT: case BT_UCHAR: return v->u; case BT_DOUBLE: case BT_FLOAT: return uint64_t(v->d); default: return uint64_t(v->i); }
}
bool Bi4Item::get_bool(const std::string& n,bool def)const{ const Bi4Value* v=find(n); return v? (v->i!=0||v->u!=0): def; }
bool Bi4Item::get_double3(const std::string& n,double out[3])const{
  const Bi4Value* v=find(n); if(!v || (v->type!=BT_DOUBLE3 && v->type!=BT_FLOAT3))return false;
  out[0]=v->d3[0]; out[1]=v->d3[1]; out[2]=v->d3[2]; return true;
}
const Bi4Array* Bi4Item::array(const std::string& n)const{ for(auto& a:arrays)if(a.name==n)return &a; return nullptr; }
Bi4Item* Bi4Item::item(const std::string& n)const{ for(auto& it:items)if(it->name==n)return it.get(); return nullptr; }
Bi4Item* Bi4Item::add_item(const std::string& n){ items.emplace_back(new Bi4Item(n)); return items.back().get(); }
void Bi4Item::add_array_ext(const std::string& n,int type,uint64_t count,const void* p){ arrays.emplace_back(); Bi4Array& a=arrays.back(); a.name=n; a.type=type; a.count=count; a.ext=p; }

// ---- writer ---------------------------------------------------------------------------------------
namespace {
struct Buf {
  std::vector<char> b;
  void raw(const void* p,size_t n){ const char* c=(const char*)p; b.insert(b.end(),c,c+n); }
  void u32(uint32_t v){ raw(&v,4); }
  void i32(int32_t v){ raw(&v,4); }
  void u64(uint64_t v){ raw(&v,8); }
  void str(const std::string& s){ u32(uint32_t(s.size())); raw(s.data(),s.size()); }
};
void put_value(Buf& o,const Bi4Value& v){
  o.str(v.name); o.i32(v.type);
  switch(v.type){
    case BT_TEXT: o.str(v.text); break;
    case BT_BOOL: o.i32(v.i!=0? 1: 0); break;
    case BT_CHAR: { int8_t x=int8_t(v.i); o.raw(&x,1); } break;
    case BT_UCHAR: { uint8_t x=uint8_t(v.u); o.raw(&x,1); } break;
    case BT_SHORT: { int16_t x=int16_t(v.i); o.raw(&x,2); } break;
    case BT_USHORT: { uint16_t x=uint16_t(v.u); o.raw(&x,2); } break;
    case BT_INT: o.i32(int32_t(v.i)); break;
    case BT_UINT: o.u32(uint32_t(v.u)); break;
    case BT_LLONG: { int64_t x=v.i; o.raw(&x,8); } break;
    case BT_ULLONG: o.u64(v.u); break;
    case BT_FLOAT: { float x=float(v.d); o.raw(&x,4); } break;
    case BT_DOUBLE: o.raw(&v.d,8); break;
    case BT_INT3: { int32_t x[3]={int32_t(v.d3[0]),int32_t(v.d3[1]),int32_t(v.d3[2])}; o.raw(x,12); } break;
    case BT_UINT3: { uint32_t x[3]={uint32_t(v.d3[0]),uint32_t(v.d3[1]),uint32_t(v.d3[2])}; o.raw(x,12); } break;
    case BT_FLOAT3: { float x[3]={float(v.d3[0]),float(v.d3[1]),float(v.d3[2])}; o.raw(x,12); } break;
    case BT_DOUBLE3: o.raw(v.d3,24); break;
    default: throw std::runtime_error("bi4: invalid value type");
  }
}
void write_item(FILE* f,const Bi4Item& it,bool si64){
  Buf vals; vals.str("\nVALUES"); vals.u32(uint32_t(it.values.size()));
  for(auto& v:it.values)put_value(vals,v);
  Buf base; base.str("\nITEM\n"); base.str(it.name); base.i32(it.hide?1:0); base.i32(it.hidevalues?1:0);
  base.str(it.fmtfloat); base.str(it.fmtdouble); base.u32(uint32_t(it.arrays.size())); base.u32(uint32_t(it.items.size())); base.u32(uint32_t(vals.b.size()));
  uint32_t sz=uint32_t(base.b.size());
  fwrite(&sz,4,1,f); fwrite(base.b.data(),1,base.b.size(),f); fwrite(vals.b.data(),1,vals.b.size(),f);
  for(auto& a:it.arrays){
    const uint64_t bytes=a.count*bi4_type_size(a.type);
    Buf ab; ab.str("\nARRAY"); ab.str(a.name); ab.i32(0); ab.i32(a.type);
    if(si64){ ab.u64(a.count); ab.u64(bytes); } else { ab.u32(uint32_t(a.count)); ab.u32(uint32_t(bytes)); }
    uint32_t asz=uint32_t(ab.b.size());
    fwrite(&asz,4,1,f); fwrite(ab.b.data(),1,ab.b.size(),f);
    if(bytes)fwrite(a.ptr(),1,size_t(bytes),f);
  }
  for(auto& ch:it.items)write_item(f,*ch,si64);
}
uint64_t item_bytes(const Bi4Item& it){
  uint64_t s=256;
  for(auto& a:it.arrays)s+=a.count*bi4_type_size(a.type)+64;
  for(auto& ch:it.items)s+=item_bytes(*ch);
  return s;
}
}

void bi4_save(const std::string& file,const Bi4Item& root,const std::string& filecode){
  const bool si64=(item_bytes(root)>=4228102018ull);
  FILE* f=fopen(file.c_str(),"wb");
  if(!f)throw std::runtime_error("bi4: cannot write "+file);
  char head[64]; memset(head,0,64);
  std::string titu="#FileJBD "+filecode;
  for(size_t c=0;c<58;c++)head[c]=(c<titu.size()? titu[c]: ' ');
  head[58]='\n'; head[60]=char(si64? 10: 0); head[61]=char(si64? 1: 0);
  fwrite(head,1,64,f);
  write_item(f,root,si64);
  if(fclose(f)!=0)throw std::runtime_error("bi4: write failure "+file);
}

// ---- reader ---------------------------------------------------------------------------------------
namespace {
struct Rd {
  const std::vector<char>& b; size_t p; bool si64;
  Rd(const std::vector<char>& v,bool s):b(v),p(0),si64(s){}
  void need(size_t n){ if(p+n>b.size())throw std::runtime_error("bi4: truncated file"); }
  uint32_t u32(){ need(4); uint32_t v; memcpy(&v,&b[p],4); p+=4; return v; }
  int32_t i32(){ need(4); int32_t v; memcpy(&v,&b[p],4); p+=4; return v; }
  uint64_t u64(){ need(8); uint64_t v; memcpy(&v,&b[p],8); p+=8; return v; }
  std::string str(){ uint32_t n=u32(); need(n); std::string s(&b[p],&b[p]+n); p+=n; return s; }
  Bi4Value value(){
    Bi4Value v; v.name=str(); v.type=i32(); v.i=0; v.u=0; v.d=0; v.d3[0]=v.d3[1]=v.d3[2]=0;
    switch(v.type){
      case BT_TEXT: v.text=str(); break;
      case BT_BOOL: v.i=i32(); break;
      case BT_CHAR: { need(1); v.i=int8_t(b[p]); p+=1; } break;
      case BT_UCHAR: { need(1); v.u=uint8_t(b[p]); p+=1; } break;
      case BT_SHORT: { need(2); int16_t x; memcpy(&x,&b[p],2); p+=2; v.i=x; } break;
      case BT_USHORT: { need(2); uint16_t x; memcpy(&x,&b[p],2); p+=2; v.u=x; } break;
      case BT_INT: v.i=i32(); break;
      case BT_UINT: v.u=u32(); break;
      case BT_LLONG: v.i=int64_t(u64()); break;
      case BT_ULLONG: v.u=u64(); break;
      case BT_FLOAT: { need(4); float x; memcpy(&x,&b[p],4); p+=4; v.d=x; } break;
      case BT_DOUBLE: { need(8); memcpy(&v.d,&b[p],8); p+=8; } break;
      case BT_INT3: { need(12); int32_t x[3]; memcpy(x,&b[p],12); p+=12; for(int k=0;k<3;k++)v.d3[k]=x[k]; } break;
      case BT_UINT3: { need(12); uint32_t x[3]; memcpy(x,&b[p],12); p+=12; for(int k=0;k<3;k++)v.d3[k]=x[k]; } break;
      case BT_FLOAT3: { need(12); float x[3]; memcpy(x,&b[p],12); p+=12; for(int k=0;k<3;k++)v.d3[k]=x[k]; } break;
      case BT_DOUBLE3: { need(24); memcpy(v.d3,&b[p],24); p+=24; } break;
      default: throw std::runtime_error("bi4: invalid value type in file");
    }
    return v;
  }
  std::unique_ptr<Bi4Item> item(){
    u32();
    if(str()!="\nITEM\n")throw std::runtime_error("bi4: ITEM code expected");
    std::unique_ptr<Bi4Item> it(new Bi4Item(str()));
    it->hide=(i32()!=0); it->hidevalues=(i32()!=0); it->fmtfloat=str(); it->fmtdouble=str();
    const uint32_t narrays=u32(), nitems=u32(), sizevalues=u32();
    if(sizevalues){
      const size_t end=p+sizevalues;
      if(str()!="\nVALUES")throw std::runtime_error("bi4: VALUES code expected");
      const uint32_t n=u32();
      for(uint32_t c=0;c<n;c++)it->values.push_back(value());
      p=end;
    }
    for(uint32_t c=0;c<narrays;c++){
      u32();
      if(str()!="\nARRAY")throw std::runtime_error("bi4: ARRAY code expected");
      Bi4Array a; a.name=str(); i32(); a.type=i32();
      uint64_t bytes;
      if(si64){ a.count=u64(); bytes=u64(); } else { a.count=u32(); bytes=u32(); }
      need(size_t(bytes));
      a.data.assign(&b[p],&b[p]+bytes); p+=size_t(bytes);
      it->arrays.push_back(std::move(a));
    }
    for(uint32_t c=0;c<nitems;c++)it->items.push_back(item());
    return it;
  }
};
}

std::unique_ptr<Bi4Item> bi4_load(const std::string& file){
  std::ifstream f(file.c_str(),std::ios::binary);
  if(!f)throw std::runtime_error("bi4: cannot open "+file);
  f.seekg(0,std::ios::end); const size_t n=size_t(f.tellg()); f.seekg(0,std::ios::beg);
  if(n<64)throw std::runtime_error("bi4: file too small "+file);
  char head[64]; f.read(head,64);
  if(strncmp(head,"#FileJBD ",9)!=0)throw std::runtime_error("bi4: not a JBinaryData file "+file);
  const bool si64=(head[61]!=0);
  const int bo=int((unsigned char)head[60])-(si64? 10: 0);
  if(bo!=0)throw std::runtime_error("bi4: byte order not supported "+file);
  std::vector<char> buf(n-64);
  f.read(buf.data(),n-64);
  if(!f)throw std::runtime_error("bi4: read failure "+file);
  Rd r(buf,si64);
  return r.item();
}

std::vector<std::unique_ptr<Bi4Item>> bi4_load_list(const std::string& file){
  std::ifstream f(file.c_str(),std::ios::binary);
  if(!f)throw std::runtime_error("bi4: cannot open "+file);
  f.seekg(0,std::ios::end); const size_t n=size_t(f.tellg()); f.seekg(0,std::ios::beg);
  if(n<64)throw std::runtime_error("bi4: file too small "+file);
  char head[64]; f.read(head,64);
  if(strncmp(head,"#FileJBD ",9)!=0)throw std::runtime_error("bi4: not a JBinaryData file "+file);
  const bool si64=(head[61]!=0);
  std::vector<char> buf(n-64);
  f.read(buf.data(),n-64);
  if(!f)throw std::runtime_error("bi4: read failure "+file);
  Rd r(buf,si64);
  std::vector<std::unique_ptr<Bi4Item>> out;
  while(r.p<buf.size())out.push_back(r.item());
  return out;
}

void bi4_save_listapp(const std::string& file,const std::string& filecode,const Bi4Item& head,const Bi4Item* item,bool truncate){
  bool exists=false;
  if(!truncate){ FILE* t=fopen(file.c_str(),"rb"); if(t){ fseek(t,0,SEEK_END); exists=(ftell(t)>=64); fclose(t); } }
  FILE* f=fopen(file.c_str(),exists? "ab": "wb");
  if(!f)throw std::runtime_error("bi4: cannot write "+file);
  if(!exists){
    char hd[64]; memset(hd,0,64);
    std::string titu="#FileJBD "+filecode;
    for(size_t c=0;c<58;c++)hd[c]=(c<titu.size()? titu[c]: ' ');
    hd[58]='\n';
    fwrite(hd,1,64,f);
    write_item(f,head,false);
  }
  if(item)write_item(f,*item,false);
  if(fclose(f)!=0)throw std::runtime_error("bi4: write failure "+file);
}

void bi4_add_array_owned(Bi4Item& it,const std::string& n,int type,uint64_t count,const void* p){
  it.arrays.emplace_back(); Bi4Array& a=it.arrays.back(); a.name=n; a.type=type; a.count=count;
  const size_t bytes=size_t(count)*bi4_type_size(type);
  a.data.assign((const char*)p,(const char*)p+bytes);
}

} // namespace b2s
