// Test infrastructure: the pair-lane HalfCheetah algorithm (envpool_b200/csrc/mujoco_pair.cuh,
// the SAME source the CUDA kernel compiles) built as plain C++.  Two host threads play the two
// lanes of an env and meet at every exchange (the device's __shfl_xor), so the CPU test suite
// can check the lane-split algebra against the oracle without a GPU.
#define HCP_HOST 1
#include <barrier>
#include <cstring>
#include <thread>

#include "../../envpool_b200/csrc/mujoco_pair.cuh"

namespace epb {
namespace hcp {

struct Chan {
  std::barrier<> bar{2};
  double slot[2];
};

double host_xch(void* chan, int side, double v) {
  Chan* ch = static_cast<Chan*>(chan);
  ch->slot[side] = v;
  ch->bar.arrive_and_wait();
  const double r = ch->slot[side ^ 1];
  ch->bar.arrive_and_wait();
  return r;
}

}  // namespace hcp
}  // namespace epb

using namespace epb;

// nsub mj_steps of one env from (q, v, warm)[9] with ctrl[6]; state is updated in place.
// ks = rows per lane held in the "shared" buffer (smaller values exercise the overflow path).
// Returns 0, or 1 if the duplicated root state of the two lanes is not bit-identical.
extern "C" int hc_pair_host_step(const void* model_blob, double* q, double* v, double* w,
                                 const double* ctrl, int nsub, int ks) {
  hcm::HcModel cm;
  std::memcpy(&cm, model_blob, sizeof(cm));
  hcm::LegModel lm[2];
  hcm::leg_model_of(cm, 0, &lm[0]);
  hcm::leg_model_of(cm, 1, &lm[1]);
  hcp::Chan chan;
  hcp::PairState st[2];
  auto lane = [&](int side) {
    hcp::PairState& s = st[side];
    const int l0 = 3 + 3 * side;
    for (int i = 0; i < 3; ++i) {
      s.qr[i] = q[i]; s.vr[i] = v[i]; s.wr[i] = w[i];
      s.ql[i] = q[l0 + i]; s.vl[i] = v[l0 + i]; s.wl[i] = w[l0 + i];
      s.ctrl[i] = ctrl[3 * side + i];
    }
    double srow[hcp::MAXR * hcp::NF], ovf[hcp::MAXR * hcp::NF];
    hcp::Ctx c;
    c.side = side; c.pm = 0; c.chan = &chan;
    c.srow = srow; c.ks = ks; c.ovf = ovf;
    for (int k = 0; k < nsub; ++k) hcp::pair_substep(c, cm, lm[side], s);
  };
  std::thread t1(lane, 1);
  lane(0);
  t1.join();
  int bad = 0;
  for (int i = 0; i < 3; ++i) {
    bad |= std::memcmp(&st[0].qr[i], &st[1].qr[i], 8) != 0;
    bad |= std::memcmp(&st[0].vr[i], &st[1].vr[i], 8) != 0;
    bad |= std::memcmp(&st[0].wr[i], &st[1].wr[i], 8) != 0;
    q[i] = st[0].qr[i]; v[i] = st[0].vr[i]; w[i] = st[0].wr[i];
    q[3 + i] = st[0].ql[i]; v[3 + i] = st[0].vl[i]; w[3 + i] = st[0].wl[i];
    q[6 + i] = st[1].ql[i]; v[6 + i] = st[1].vl[i]; w[6 + i] = st[1].wl[i];
  }
  return bad;
}
