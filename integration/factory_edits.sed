# integration/factory_edits.sed -- the TWO factory lines a maintainer changes in the reference tree
# (plus the includes they need).  oracle/Makefile applies this to temporary copies of
#   src/loss/loss.cc          :15   loss = new FMLoss();        ->  loss = new GpuFMLoss();
#   src/sgd/sgd_learner.cc    :233  auto updater = new SGDUpdater();  ->  auto updater = new GpuSGDUpdater();
# in its build directory (oracle/_ref/obj); the reference tree itself is read-only and nothing of it is copied
# into this repository.
s|loss = new FMLoss();|loss = new GpuFMLoss();|
s|auto updater = new SGDUpdater();|auto updater = new GpuSGDUpdater();|
s|#include "./fm_loss.h"|#include "./fm_loss.h"\n#include "gpu_fm_loss.h"|
s|#include "./sgd_learner.h"|#include "./sgd_learner.h"\n#include "gpu_sgd_updater.h"|
