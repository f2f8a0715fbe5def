// ref_dbow2_wrap.cpp -- C entry points around the REFERENCE's own DBoW2 (compiled unmodified from
// /root/reference/code/Thirdparty/DBoW2 by `make -C oracle ref`, against oracle/ref_shim's stand-in for
// opencv2/core/core.hpp).  Test infrastructure only: tests/test_ref_dbow2.py uses it to pin the oracle's restatement
// (orc_bow_transform, dist256) against the reference's code itself.  Nothing in the product links or loads this.
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "DBoW2/FORB.h"
#include "DBoW2/TemplatedVocabulary.h"

typedef DBoW2::TemplatedVocabulary<DBoW2::FORB::TDescriptor, DBoW2::FORB> RefVocabulary;  // = ORB_SLAM2::ORBVocabulary

extern "C" {

void* ref_vocab_load(const char* path) {
  RefVocabulary* v = new RefVocabulary();
  if (!v->loadFromBinaryFile(path)) {
    delete v;
    return nullptr;
  }
  return v;
}

void ref_vocab_destroy(void* v) { delete static_cast<RefVocabulary*>(v); }

void ref_vocab_info(void* vp, int32_t* k, int32_t* L, int32_t* n_words) {
  RefVocabulary* v = static_cast<RefVocabulary*>(vp);
  *k = v->getBranchingFactor();
  *L = v->getDepthLevels();
  *n_words = (int32_t)v->size();
}

// mpORBvocabulary->transform(vCurrentDesc, mBowVec, mFeatVec, levelsup), flattened in the containers' own order
int ref_bow_transform(void* vp, const uint8_t* desc, int n, int levelsup, uint32_t* word_ids, double* word_values,
                      uint32_t* node_ids, int32_t* offsets, uint32_t* feats, int32_t* n_nodes) {
  RefVocabulary* v = static_cast<RefVocabulary*>(vp);
  std::vector<cv::Mat> features;
  features.reserve(n);
  for (int i = 0; i < n; i++) {
    cv::Mat d(1, 32, CV_8U);
    std::memcpy(d.data, desc + (size_t)i * 32, 32);
    features.push_back(d);
  }
  DBoW2::BowVector bv;
  DBoW2::FeatureVector fv;
  v->transform(features, bv, fv, levelsup);
  int nw = 0;
  for (DBoW2::BowVector::const_iterator it = bv.begin(); it != bv.end(); ++it) {
    word_ids[nw] = it->first;
    word_values[nw] = it->second;
    nw++;
  }
  int nn = 0, nf = 0;
  offsets[0] = 0;
  for (DBoW2::FeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it) {
    node_ids[nn] = it->first;
    for (size_t j = 0; j < it->second.size(); j++) feats[nf++] = it->second[j];
    offsets[++nn] = nf;
  }
  *n_nodes = nn;
  return nw;
}

// FORB::distance (FORB.cpp:82-102), the same bit-count ORBmatcher::DescriptorDistance uses (ORBmatcher.cc:1511-1525)
int ref_forb_distance(const uint8_t* a, const uint8_t* b) {
  cv::Mat ma(1, 32, CV_8U), mb(1, 32, CV_8U);
  std::memcpy(ma.data, a, 32);
  std::memcpy(mb.data, b, 32);
  return DBoW2::FORB::distance(ma, mb);
}

// L1 score of two bag-of-words vectors (ScoringObject.cpp L1Scoring::score), as KeyFrameDatabase uses it
double ref_score(void* vp, const uint32_t* ids1, const double* vals1, int n1, const uint32_t* ids2, const double* vals2, int n2) {
  RefVocabulary* v = static_cast<RefVocabulary*>(vp);
  DBoW2::BowVector a, b;
  for (int i = 0; i < n1; i++) a.addWeight(ids1[i], vals1[i]);
  for (int i = 0; i < n2; i++) b.addWeight(ids2[i], vals2[i]);
  return v->score(a, b);
}

}  // extern "C"
