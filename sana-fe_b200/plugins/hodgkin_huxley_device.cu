// hodgkin_huxley_device.cu — the reference's Hodgkin-Huxley soma plugin (plugins/hodgkin_huxley.cpp) as an OUT-OF-TREE
// device model for the B200 engine: the worked example of include/sfe_device_model.h.
//
//   reference                                   here
//   class HodgkinHuxley : SomaUnit              struct HodgkinHuxley (a functor, no virtuals)
//   register_attributes({"m","n","h","current"}) state_name / param_name
//   set_attribute_neuron                        filled by the engine's lowering from those names
//   update(neuron_address, current_in, ts)      __device__ update(state, params, has_in, in, ts)   (:116-170)
//   reset()                                     kResetMask: V, m, n, h -> 0 (:70-86; I is kept)
//   get_potential()                             kPotentialState = 0 (V)
//   extern "C" create_hodgkin_huxley()          SFE_DEVICE_SOMA_MODEL(hodgkin_huxley, HodgkinHuxley)
//
// The plugin keeps one neuron of state per hardware unit and ignores current_in (as the reference does): map one
// neuron per unit. Built by sana-fe_b200/Makefile into sanafe_b200/plugins/libhodgkin_huxley_b200.so.
#include "sfe_device_model.h"

struct HodgkinHuxley
{
    static constexpr int kState = 4, kParams = 1;   // V, m, n, h | I
    static constexpr unsigned kResetMask = 0xF;
    static constexpr int kPotentialState = 0;
    static const char *state_name(const int w)
    {
        static const char *names[] = {"", "m", "n", "h"}; // V has no attribute: it starts at 0
        return names[w];
    }
    static const char *param_name(int) { return "current"; }
    static double state_init(int) { return 0.0; }
    static double param_init(int) { return 0.0; }

    __device__ static int update(const sfe_model_words<double> s, const sfe_model_words<const double> p, bool /*has_in*/,
            double /*in*/, long long /*timestep*/)
    {
        // system constants of the plugin (plugins/hodgkin_huxley.cpp:27-35)
        const double C_m = 10.0, g_Na = 1200.0, g_K = 360.0, g_L = 3.0, V_Na = 50.0, V_K = -77.0, V_L = 54.387, dt = 0.1;
        double V = s[0], m = s[1], n = s[2], h = s[3];
        const double I = p[0];
        const double alpha_n = (0.01 * (V + 55)) / (1 - exp(-0.1 * (V + 55)));
        const double alpha_m = (0.1 * (V + 40)) / (1 - exp(-0.1 * (V + 40)));
        const double alpha_h = 0.07 * exp(-0.05 * (V + 65));
        const double beta_n = 0.125 * exp(-0.01125 * (V + 55));
        const double beta_m = 4 * exp(-0.05556 * (V + 65));
        const double beta_h = 1 / (1 + exp(-0.1 * (V + 35)));
        const double tau_n = 1 / (alpha_n + beta_n), tau_m = 1 / (alpha_m + beta_m), tau_h = 1 / (alpha_h + beta_h);
        const double pm = alpha_m / (alpha_m + beta_m), pn = alpha_n / (alpha_n + beta_n), ph = alpha_h / (alpha_h + beta_h);
        const double n4 = pow(n, 4.0), m3 = pow(m, 3.0);
        const double denominator = g_L + g_K * n4 + g_Na * (m3 * h);
        const double tau_V = C_m / denominator;
        const double Vinf = ((g_L) * V_L + g_K * n4 * V_K + g_Na * m3 * h * V_Na + I) / denominator;
        const double prev_V = V;
        V = Vinf + (V - Vinf) * exp(-1 * dt / tau_V);
        m = pm + (m - pm) * exp(-1 * dt / tau_m);
        n = pn + (n - pn) * exp(-1 * dt / tau_n);
        h = ph + (h - ph) * exp(-1 * dt / tau_h);
        s[0] = V;
        s[1] = m;
        s[2] = n;
        s[3] = h;
        return ((prev_V < 25) && (V > 25)) ? SFE_MODEL_FIRED : SFE_MODEL_UPDATED;
    }
};

SFE_DEVICE_SOMA_MODEL(hodgkin_huxley, HodgkinHuxley)
