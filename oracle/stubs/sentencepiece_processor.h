// Stub for the third-party sentencepiece header that op/encode.h includes.
// TEST INFRASTRUCTURE ONLY (oracle build). The tokenizer is out of scope:
// op::SPELayer is constructed unconditionally by LlamaModel::create_nonparam_layers, so the
// type must exist and Load() must succeed; nothing on the forward path calls it.
#pragma once
#include <string>
#include <vector>
namespace sentencepiece {
struct Status {
    bool ok() const { return true; }
    std::string ToString() const { return std::string(); }
};
class SentencePieceProcessor {
public:
    Status Load(const std::string&) { return Status(); }
    Status Encode(const std::string&, std::vector<int>* ids) const { if (ids) ids->clear(); return Status(); }
    Status Decode(const std::vector<int>&, std::string* text) const { if (text) text->clear(); return Status(); }
    int GetPieceSize() const { return 0; }
};
}  // namespace sentencepiece
