/**
 * integration/gpu_fm_loss.h -- the Loss a difacto maintainer drops into the REFERENCE tree (next to
 * src/loss/fm_loss.h): FMLoss::Predict / CalcGrad (src/loss/fm_loss.h:56-119,136-199) and Loss::Evaluate
 * (include/difacto/loss.h:57-66) on the B200 engine through dfb_predict / dfb_calc_grad / dfb_evaluate.
 * Compiled against the unmodified reference headers by oracle/Makefile (target ref_gpu).
 * Loss::Create (src/loss/loss.cc:12-26) has no access to the updater, so the loss owns a small engine of its own:
 * these calls only use the engine's workspaces, never its table.
 */
#ifndef INTEGRATION_GPU_FM_LOSS_H_
#define INTEGRATION_GPU_FM_LOSS_H_
#include <string>
#include <vector>
#include "difacto/loss.h"
#include "loss/fm_loss.h"
#include "difacto_b200.h"
namespace difacto {

class GpuFMLoss : public Loss {
 public:
  GpuFMLoss() {}
  virtual ~GpuFMLoss() { if (h_) dfb_destroy(h_); }

  KWArgs Init(const KWArgs& kwargs) override {
    KWArgs remain = param_.InitAllowUnknown(kwargs);       // FMLossParam: V_dim (fm_loss.h:19-27)
    const std::string vd = std::to_string(param_.V_dim);
    const char* k[] = {"V_dim", "table_capacity"};
    const char* v[] = {vd.c_str(), "1024"};
    CHECK_EQ(dfb_create(k, v, 2, &h_), 0) << dfb_last_error(nullptr);
    return remain;
  }

  void Predict(const dmlc::RowBlock<unsigned>& data, const std::vector<SArray<char>>& param,
               SArray<real_t>* pred) override {
    CHECK_EQ(param.size(), 3);
    const SArray<real_t> weights(param[0]);
    const SArray<int> w_pos(param[1]), V_pos(param[2]);
    CHECK_EQ(data.offset[0], 0);
    // w_pos / V_pos are empty when V_dim == 0 (SGDLearner::GetPos on empty lens): direct indexing
    CHECK_EQ(dfb_predict(h_, data.size, reinterpret_cast<const uint64_t*>(data.offset), data.index, data.value,
                         weights.data(), weights.size(), w_pos.empty() ? nullptr : w_pos.data(),
                         V_pos.empty() ? nullptr : V_pos.data(), w_pos.size(), pred->data()), 0)
        << dfb_last_error(h_);
  }

  real_t Evaluate(dmlc::real_t const* label, const SArray<real_t>& pred) const override {
    float objv = 0;
    CHECK_EQ(dfb_evaluate(h_, label, pred.data(), pred.size(), &objv), 0) << dfb_last_error(h_);
    return objv;
  }

  void CalcGrad(const dmlc::RowBlock<unsigned>& data, const std::vector<SArray<char>>& param,
                SArray<real_t>* grad) override {
    CHECK_EQ(param.size(), 4);
    const SArray<real_t> weights(param[0]), pred(param[3]);
    const SArray<int> w_pos(param[1]), V_pos(param[2]);
    CHECK_EQ(dfb_calc_grad(h_, data.size, reinterpret_cast<const uint64_t*>(data.offset), data.index, data.value,
                           data.label, weights.data(), weights.size(), w_pos.empty() ? nullptr : w_pos.data(),
                           V_pos.empty() ? nullptr : V_pos.data(), w_pos.size(), pred.data(), grad->data()), 0)
        << dfb_last_error(h_);
  }

 private:
  FMLossParam param_;
  dfb_handle h_ = nullptr;
};

}  // namespace difacto
#endif  // INTEGRATION_GPU_FM_LOSS_H_
