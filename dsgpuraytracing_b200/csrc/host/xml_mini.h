// xml_mini.h -- a small DOM parser for the subset of XML that COLLADA scene files use (elements, attributes,
// character data, comments, <?...?> and <!...> declarations, the five predefined entities).  It stands in for the
// tinyxml2 calls the reference's ColladaParser makes (FirstChildElement / NextSiblingElement / Attribute / GetText,
// reference src/collada/collada.cpp); GetText() has tinyxml2's meaning: the character data that immediately follows
// the start tag (first child text node), or null.
#pragma once
#include <cstring>
#include <memory>
#include <string>
#include <utility>
#include <vector>

namespace dsrt_host {

struct XmlElement {
  std::string name;
  std::vector<std::pair<std::string, std::string>> attrs;
  std::string text;            // first text node (before the first child element)
  bool has_text = false;
  std::vector<std::unique_ptr<XmlElement>> children;
  XmlElement* parent = nullptr;
  int index_in_parent = 0;

  const char* Attribute(const char* n) const {
    for (const auto& a : attrs) if (a.first == n) return a.second.c_str();
    return nullptr;
  }
  int IntAttribute(const char* n) const { const char* v = Attribute(n); return v ? atoi(v) : 0; }
  const char* GetText() const { return has_text ? text.c_str() : nullptr; }
  const char* Name() const { return name.c_str(); }
  XmlElement* FirstChildElement(const char* n = nullptr) const {
    for (const auto& c : children) if (!n || c->name == n) return c.get();
    return nullptr;
  }
  XmlElement* NextSiblingElement(const char* n = nullptr) const {
    if (!parent) return nullptr;
    for (size_t i = (size_t)index_in_parent + 1; i < parent->children.size(); i++)
      if (!n || parent->children[i]->name == n) return parent->children[i].get();
    return nullptr;
  }
};

class XmlDocument {
 public:
  // returns false (and sets error()) on malformed input
  bool Parse(const std::string& s) {
    src_ = &s; pos_ = 0; err_.clear();
    root_.reset(new XmlElement());
    root_->name = "#document";
    XmlElement* cur = root_.get();
    const size_t n = s.size();
    bool seen_child = false;   // per current element: text only counts before the first child
    std::vector<bool> seen_stack;
    while (pos_ < n) {
      if (s[pos_] != '<') {
        size_t e = s.find('<', pos_);
        if (e == std::string::npos) e = n;
        if (cur != root_.get() && !seen_child && !cur->has_text) {
          std::string t = decode(s.substr(pos_, e - pos_));
          bool ws = true;
          for (char c : t) if (!isspace((unsigned char)c)) { ws = false; break; }
          if (!ws) { cur->text = t; cur->has_text = true; }
        }
        pos_ = e;
        continue;
      }
      if (s.compare(pos_, 4, "<!--") == 0) {
        size_t e = s.find("-->", pos_ + 4);
        if (e == std::string::npos) return fail("unterminated comment");
        pos_ = e + 3;
        continue;
      }
      if (s.compare(pos_, 2, "<?") == 0) {
        size_t e = s.find("?>", pos_ + 2);
        if (e == std::string::npos) return fail("unterminated declaration");
        pos_ = e + 2;
        continue;
      }
      if (s.compare(pos_, 9, "<![CDATA[") == 0) {
        size_t e = s.find("]]>", pos_ + 9);
        if (e == std::string::npos) return fail("unterminated CDATA");
        if (cur != root_.get() && !seen_child && !cur->has_text) { cur->text = s.substr(pos_ + 9, e - pos_ - 9); cur->has_text = true; }
        pos_ = e + 3;
        continue;
      }
      if (s.compare(pos_, 2, "<!") == 0) {
        size_t e = s.find('>', pos_ + 2);
        if (e == std::string::npos) return fail("unterminated <! section");
        pos_ = e + 1;
        continue;
      }
      if (s.compare(pos_, 2, "</") == 0) {
        size_t e = s.find('>', pos_ + 2);
        if (e == std::string::npos) return fail("unterminated end tag");
        std::string nm = trim(s.substr(pos_ + 2, e - pos_ - 2));
        if (cur == root_.get() || nm != cur->name) return fail("mismatched end tag </" + nm + ">");
        cur = cur->parent;
        seen_child = seen_stack.back(); seen_stack.pop_back();
        pos_ = e + 1;
        continue;
      }
      // start tag
      size_t p = pos_ + 1;
      size_t q = p;
      while (q < n && !isspace((unsigned char)s[q]) && s[q] != '>' && s[q] != '/') q++;
      std::unique_ptr<XmlElement> el(new XmlElement());
      el->name = s.substr(p, q - p);
      if (el->name.empty()) return fail("empty element name");
      bool self_close = false;
      while (true) {
        while (q < n && isspace((unsigned char)s[q])) q++;
        if (q >= n) return fail("unterminated start tag");
        if (s[q] == '>') { q++; break; }
        if (s[q] == '/') { if (q + 1 < n && s[q + 1] == '>') { self_close = true; q += 2; break; } return fail("stray '/'"); }
        size_t a0 = q;
        while (q < n && s[q] != '=' && !isspace((unsigned char)s[q]) && s[q] != '>') q++;
        std::string an = s.substr(a0, q - a0);
        while (q < n && isspace((unsigned char)s[q])) q++;
        if (q >= n || s[q] != '=') return fail("attribute without value: " + an);
        q++;
        while (q < n && isspace((unsigned char)s[q])) q++;
        if (q >= n || (s[q] != '"' && s[q] != '\'')) return fail("unquoted attribute value: " + an);
        const char quote = s[q++];
        size_t v0 = q;
        while (q < n && s[q] != quote) q++;
        if (q >= n) return fail("unterminated attribute value");
        el->attrs.emplace_back(an, decode(s.substr(v0, q - v0)));
        q++;
      }
      el->parent = cur;
      el->index_in_parent = (int)cur->children.size();
      XmlElement* raw = el.get();
      cur->children.push_back(std::move(el));
      seen_child = true;
      if (!self_close) { seen_stack.push_back(seen_child); cur = raw; seen_child = false; }
      pos_ = q;
    }
    if (cur != root_.get()) return fail("unexpected end of document inside <" + cur->name + ">");
    return true;
  }
  XmlElement* FirstChildElement(const char* n = nullptr) const { return root_ ? root_->FirstChildElement(n) : nullptr; }
  const std::string& error() const { return err_; }

 private:
  bool fail(const std::string& m) { err_ = m + " at byte " + std::to_string(pos_); return false; }
  static std::string trim(const std::string& s) {
    size_t a = 0, b = s.size();
    while (a < b && isspace((unsigned char)s[a])) a++;
    while (b > a && isspace((unsigned char)s[b - 1])) b--;
    return s.substr(a, b - a);
  }
  static std::string decode(const std::string& s) {
    if (s.find('&') == std::string::npos) return s;
    std::string o; o.reserve(s.size());
    for (size_t i = 0; i < s.size(); i++) {
      if (s[i] == '&') {
        if (s.compare(i, 4, "&lt;") == 0) { o += '<'; i += 3; continue; }
        if (s.compare(i, 4, "&gt;") == 0) { o += '>'; i += 3; continue; }
        if (s.compare(i, 5, "&amp;") == 0) { o += '&'; i += 4; continue; }
        if (s.compare(i, 6, "&quot;") == 0) { o += '"'; i += 5; continue; }
        if (s.compare(i, 6, "&apos;") == 0) { o += '\''; i += 5; continue; }
      }
      o += s[i];
    }
    return o;
  }
  const std::string* src_ = nullptr;
  size_t pos_ = 0;
  std::string err_;
  std::unique_ptr<XmlElement> root_;
};

}  // namespace dsrt_host
