/**
 * integration/ref_sgd_gpu_main.cc -- the reference's OWN SGDLearner (src/sgd/sgd_learner.{h,cc}, unmodified
 * except for the two factory lines of integration/factory_edits.sed) driving the GPU plugins: the drop-in proof.
 * TEST INFRASTRUCTURE: built by oracle/Makefile (target ref_gpu) into oracle/_ref/ref_sgd_gpu; mirrors
 * tests/cpp/sgd_learner_test.cc of the reference (kwargs from argv as key=value, one line per epoch on stdout).
 */
#include <cstdio>
#include <string>
#include "sgd/sgd_learner.h"

int main(int argc, char** argv) {
  using namespace difacto;
  KWArgs args;
  for (int i = 1; i < argc; ++i) {
    std::string a(argv[i]);
    size_t eq = a.find('=');
    if (eq == std::string::npos) { fprintf(stderr, "expected key=value, got %s\n", argv[i]); return 2; }
    args.push_back(std::make_pair(a.substr(0, eq), a.substr(eq + 1)));
  }
  SGDLearner learner;
  auto remain = learner.Init(args);
  for (const auto& r : remain) fprintf(stderr, "unrecognized keyword argument: %s = %s\n", r.first.c_str(), r.second.c_str());
  learner.AddEpochEndCallback([](int epoch, const sgd::Progress& train, const sgd::Progress& val) {
    printf("epoch %d loss %.9g penalty %.9g auc %.9g nrows %.9g val_loss %.9g val_auc %.9g\n", epoch, train.loss,
           train.penalty, train.auc, train.nrows, val.loss, val.auc);
    fflush(stdout);
  });
  learner.Run();
  return 0;
}
