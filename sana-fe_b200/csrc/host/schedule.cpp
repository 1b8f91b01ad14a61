// schedule.cpp — the "detailed" timing model on the host, fed by device records.
//
// The reference's semi-analytical NoC scheduler (src/schedule.cpp:208-620) is a
// strictly sequential priority-queue simulation over the messages of one
// timestep; it stays on the host. What the device hands over per
// step is one status byte per neuron (idle / updated / fired); every other
// message field is a load-time constant of the lowered tables, so the per-core
// message lists are rebuilt here exactly as process_neurons /
// pipeline_process_axon_out build them (src/chip.cpp:624-654, 710-736, 802-834):
// same order, same floating-point accumulation of generation delays.
#include "schedule.hpp"

#include <algorithm>
#include <cmath>
#include <limits>
#include <list>
#include <queue>
#include <stdexcept>

namespace sfe
{
namespace
{
enum Direction : uint32_t // src/schedule.hpp:31-39
{
    north = 0U,
    east = 1U,
    south = 2U,
    west = 3U,
    ndirections = 4U,
};

struct Msg
{
    double generation_delay{0.0};
    double processing_delay{0.0};
    double min_hop_delay{0.0};
    double sent{0.0}, received{0.0}, processed{0.0};
    uint32_t hops{0};
    uint32_t src_x{0}, src_y{0}, dest_x{0}, dest_y{0};
    uint32_t src_core_offset{0}, src_core{0}, dest_core{0};
    bool placeholder{true};
    bool in_noc{false};
    uint32_t slot{0}; // index of this message's trace record
};

struct BySentTime // CompareMessagesBySentTime  src/message.cpp:61-65
{
    bool operator()(const Msg &a, const Msg &b) const noexcept { return a.sent > b.sent; }
};

struct Noc // NocInfo  src/schedule.hpp:170-202
{
    size_t width, height, links_per_router;
    std::vector<double> density;
    std::vector<double> core_free;
    double mean_rx{0.0};
    long in_noc{0};

    size_t idx(size_t x, size_t y, size_t link) const { return (x * height * links_per_router) + (y * links_per_router) + link; }

    template <typename F> void walk_route(const Msg &m, F &&visit) const
    {
        // dimension-order route: x first, then y; first link = the sending core's
        // injection link (src/schedule.cpp:449-611)
        const int xi = (m.src_x < m.dest_x) ? 1 : -1;
        const int yi = (m.src_y < m.dest_y) ? 1 : -1;
        size_t prev = ndirections + m.src_core_offset;
        for (size_t x = m.src_x; x != m.dest_x; x += xi)
        {
            const size_t dir = (xi > 0) ? east : west;
            if (x == m.src_x) visit(idx(x, m.src_y, ndirections + m.src_core_offset));
            else visit(idx(x, m.src_y, dir));
            prev = dir;
        }
        for (size_t y = m.src_y; y != m.dest_y; y += yi)
        {
            const size_t dir = (yi > 0) ? north : south;
            if (m.src_x == m.dest_x && y == m.src_y) visit(idx(m.dest_x, y, ndirections + m.src_core_offset));
            else visit(idx(m.dest_x, y, prev));
            prev = dir;
        }
        if (m.src_x == m.dest_x && m.src_y == m.dest_y) visit(idx(m.dest_x, m.dest_y, ndirections + m.src_core_offset));
        else visit(idx(m.dest_x, m.dest_y, prev));
    }

    double congestion(const Msg &m) const
    {
        double flow = 0.0;
        walk_route(m, [&](size_t link) { flow += density[link]; });
        return flow;
    }

    void track(const Msg &m, const bool entering)
    {
        if (m.src_x > width || m.dest_x > width) throw std::runtime_error("Message x > NoC width");
        if (m.src_y > height || m.dest_y > height) throw std::runtime_error("Message y > NoC height");
        double adjust = 1.0 / (2.0 + static_cast<double>(m.hops));
        if (!entering) adjust *= -1.0;
        walk_route(m, [&](size_t link) { density[link] += adjust; });
        // update_rolling_averages  src/schedule.cpp:449-475
        if (entering)
        {
            mean_rx += (m.processing_delay - mean_rx) / (static_cast<double>(in_noc) + 1.0);
            in_noc++;
        }
        else
        {
            if (in_noc > 1) mean_rx += (mean_rx - m.processing_delay) / (static_cast<double>(in_noc) - 1.0);
            else mean_rx = 0.0;
            in_noc--;
        }
    }
};
struct InFlight // a message between being sent and being fully received
{
    double received;
    uint32_t dest_core;
    uint32_t seq; // arrival order (global counter: monotonic, hence also the order within a core)
    uint32_t msg;
};

struct NeuronLatency // per soma class: what a neuron adds to its core's generation delay, by status
{
    double dend, access, update, spike_out;
    bool dend_in_neuron;
};
} // namespace

struct DetailedScheduler::Scratch // reused from step to step
{
    std::vector<Msg> msgs;
    std::vector<uint32_t> core_begin, head;
    std::vector<InFlight> in_flight, due;
    std::vector<NeuronLatency> neuron_latency;
    std::vector<MessageRecord> recs;
    std::vector<const MessageRecord *> order;
    uint32_t n_classes_seen{0};
    Noc noc;
};

DetailedScheduler::~DetailedScheduler() = default;

DetailedScheduler::DetailedScheduler(const sfe_tables &t) : t_(t), scratch_(std::make_unique<Scratch>())
{
    axon_core_.resize(t.n_axons_in);
    axon_proc_.resize(t.n_axons_in);
    for (uint32_t c = 0; c < t.n_cores; ++c)
    {
        const sfe_core_desc &cd = t.cores[c];
        for (uint32_t a = 0; a < cd.axon_in_count; ++a)
        {
            const uint32_t id = cd.axon_in_begin + a;
            axon_core_[id] = c;
            // process_message  src/chip.cpp:738-764: axon-in latency, then per synapse
            // the pipeline's (0.0 + synapse) + dendrite latency, added one by one
            const sfe_axon_in &ax = t.axons_in[id];
            const sfe_cost_class &cc = t.cost_classes[ax.cost_class];
            double lat = cd.latency_axon_in;
            if (cc.per_message) lat += cc.syn_latency;
            else
                for (uint32_t s = 0; s < ax.syn_count; ++s) lat += (0.0 + cc.syn_latency) + cc.den_latency;
            axon_proc_[id] = lat;
        }
    }
}

double DetailedScheduler::run_step(
        const uint8_t *status, const bool schedule, std::vector<MessageRecord> *trace, long *next_mid)
{
    const sfe_tables &t = t_;
    const double ninf = -std::numeric_limits<double>::infinity();
    // creation order = cores ascending, in-core send order (kept in the scratch: a few hundred KB per step that
    // would otherwise be mapped and unmapped every step, which serialises the scheduler threads in the kernel)
    std::vector<MessageRecord> &recs = scratch_->recs;
    recs.clear();
    auto new_record = [&](const Msg &m, const uint32_t src_neuron, const uint32_t spikes) -> uint32_t {
        MessageRecord r;
        r.placeholder = m.placeholder;
        r.mid = m.placeholder ? -1L : (*next_mid)++;
        r.src_neuron = src_neuron;
        r.src_core = m.src_core;
        r.dest_core = m.dest_core;
        r.hops = m.hops;
        r.spikes = spikes;
        r.sent = r.received = r.processed = ninf;
        r.generation_delay = m.generation_delay;
        r.processing_delay = m.processing_delay;
        r.min_hop_delay = m.min_hop_delay;
        recs.push_back(r);
        return static_cast<uint32_t>(recs.size() - 1);
    };
    // ---- rebuild the per-sending-core message lists ------------------------------
    // (one flat array in creation order; a core's messages are contiguous, `head` walks them)
    Scratch &sc = *scratch_;
    std::vector<Msg> &msgs = sc.msgs;
    std::vector<uint32_t> &core_begin_ = sc.core_begin, &head_ = sc.head;
    msgs.clear();
    core_begin_.assign(t.n_cores + 1, 0);
    if (sc.n_classes_seen != t.n_soma_classes) // (MappedNeuron.set_attributes can add classes between runs)
    {
        sc.neuron_latency.resize(t.n_soma_classes);
        for (uint32_t k = 0; k < t.n_soma_classes; ++k)
        {
            const sfe_soma_class &cls = t.soma_classes[k];
            sc.neuron_latency[k] = {cls.dend_latency_update, cls.latency_access, cls.latency_update, cls.latency_spike_out,
                    cls.dend_in_neuron != 0};
        }
        sc.n_classes_seen = t.n_soma_classes;
    }
    const std::vector<NeuronLatency> &neuron_latency_ = sc.neuron_latency;
    for (uint32_t c = 0; c < t.n_cores; ++c)
    {
        core_begin_[c] = static_cast<uint32_t>(msgs.size());
        const sfe_core_desc &cd = t.cores[c];
        if (cd.neuron_count == 0) continue;
        const sfe_tile_desc &src_tile = t.tiles[cd.tile];
        double next_delay = 0.0;
        for (uint32_t k = 0; k < cd.neuron_count; ++k)
        {
            const uint32_t i = cd.neuron_begin + k;
            const uint8_t st = status[i];
            const NeuronLatency &nl = neuron_latency_[t.neuron_class[i]];
            // dendrite (when it runs in the neuron pipeline) then soma: access, + update, + spike out
            double lat = 0.0;
            if (nl.dend_in_neuron) lat += nl.dend;
            double soma = nl.access;
            if (st == SFE_STATUS_UPDATED || st == SFE_STATUS_FIRED) soma += nl.update;
            if (st == SFE_STATUS_FIRED) soma += nl.spike_out;
            lat += soma;
            next_delay += lat;
            if (st != SFE_STATUS_FIRED) continue;
            for (uint32_t a = t.axon_out_begin[i]; a < t.axon_out_begin[i + 1]; ++a)
            {
                const uint32_t id = t.axon_out_target[a];
                const uint32_t dc = axon_core_[id];
                const sfe_core_desc &dd = t.cores[dc];
                const sfe_tile_desc &dst_tile = t.tiles[dd.tile];
                const sfe_axon_in &ax = t.axons_in[id];
                Msg m;
                m.placeholder = false;
                m.generation_delay = next_delay + cd.latency_axon_out;
                next_delay = 0.0;
                m.processing_delay = axon_proc_[id];
                const uint32_t dx = SFE_HOP_DX(ax.hop), dy = SFE_HOP_DY(ax.hop);
                m.hops = dx + dy;
                // sim_estimate_network_costs  src/chip.cpp:1127-1169 (source tile's latencies)
                m.min_hop_delay = 0.0;
                m.min_hop_delay += static_cast<double>(dx) * (SFE_HOP_EAST(ax.hop) ? src_tile.latency_east : src_tile.latency_west);
                m.min_hop_delay += static_cast<double>(dy) * (SFE_HOP_NORTH(ax.hop) ? src_tile.latency_north : src_tile.latency_south);
                m.src_x = src_tile.x;
                m.src_y = src_tile.y;
                m.dest_x = dst_tile.x;
                m.dest_y = dst_tile.y;
                m.src_core_offset = cd.offset;
                m.src_core = c;
                m.dest_core = dc;
                if (trace != nullptr) m.slot = new_record(m, i, ax.syn_count);
                msgs.push_back(m);
            }
        }
        if (next_delay != 0.0)
        {
            Msg m; // placeholder: processing that sends nothing (src/chip.cpp:640-652)
            m.generation_delay = next_delay;
            m.src_x = src_tile.x;
            m.src_y = src_tile.y;
            m.src_core_offset = cd.offset;
            m.src_core = c;
            if (trace != nullptr) m.slot = new_record(m, cd.neuron_begin + cd.neuron_count - 1, 0u);
            msgs.push_back(m);
        }
    }
    core_begin_[t.n_cores] = static_cast<uint32_t>(msgs.size());
    auto finish_trace = [&]() {
        if (trace == nullptr) return;
        // sim_sort_and_record_messages: the same std::sort on the same sequence with the same
        // comparator (placeholders compare equal, so their order is whatever the sort leaves)
        std::vector<const MessageRecord *> &order = scratch_->order;
        order.clear();
        order.reserve(recs.size());
        trace->reserve(trace->size() + recs.size());
        for (const MessageRecord &r : recs) order.push_back(&r);
        std::sort(order.begin(), order.end(), [](const MessageRecord *a, const MessageRecord *b) {
            if (a->placeholder && b->placeholder) return a->mid < b->mid; // CompareMessagesByID  src/message.cpp:70-91
            if (a->placeholder) return false;
            if (b->placeholder) return true;
            return a->mid < b->mid;
        });
        for (const MessageRecord *r : order) trace->push_back(*r);
    };
    if (!schedule)
    {
        // schedule_messages_timestep_simple  src/schedule.cpp:84-86: no blocking, network = minimum hops
        for (MessageRecord &r : recs) r.network_delay = r.min_hop_delay;
        finish_trace();
        return 0.0;
    }

    // ---- schedule_messages_timestep_detailed  src/schedule.cpp:208-292 -----------------
    // Same event order and the same floating-point operations in the same order as the reference; only the
    // containers differ. The reference keeps a std::list of in-flight messages per destination core and, for every
    // event, walks ALL of them to find those received by then (noc_update_all_tracked_messages, :380-400). Here the
    // in-flight messages sit in a min-heap by received time: the ones that are due are popped and then retired in
    // the order the reference's walk would meet them (destination core ascending, arrival order within a core).
    Noc &noc = sc.noc;
    noc.width = t.noc_width;
    noc.height = t.noc_height;
    noc.links_per_router = t.max_cores_per_tile + ndirections;
    noc.core_free.assign(t.n_cores, 0.0);
    noc.density.assign(static_cast<size_t>(t.noc_width) * t.noc_height * noc.links_per_router, 0.0);
    noc.mean_rx = 0.0;
    noc.in_noc = 0;
    head_.assign(core_begin_.begin(), core_begin_.end() - 1);
    // event queue: std::priority_queue over (sent time, message) with the reference's comparator on the sent
    // time alone — the heap's tie behaviour depends only on the sequence of pushes and pops, which is the same
    struct Event
    {
        double sent;
        uint32_t msg;
    };
    struct LaterFirst
    {
        bool operator()(const Event &a, const Event &b) const noexcept { return a.sent > b.sent; }
    };
    std::priority_queue<Event, std::vector<Event>, LaterFirst> pq;
    for (uint32_t c = 0; c < t.n_cores; ++c)
    {
        if (head_[c] == core_begin_[c + 1]) continue;
        Msg &m = msgs[head_[c]++];
        m.sent = m.generation_delay;
        if (trace != nullptr) recs[m.slot].sent = m.sent;
        pq.push({m.sent, static_cast<uint32_t>(&m - msgs.data())});
    }
    auto later_received = [](const InFlight &a, const InFlight &b) { return a.received > b.received; };
    std::vector<InFlight> &in_flight = sc.in_flight;
    std::vector<InFlight> &due = sc.due;
    in_flight.clear();
    uint32_t seq = 0;
    double last = 0.0;
    while (!pq.empty())
    {
        const Event ev = pq.top();
        pq.pop();
        Msg &m = msgs[ev.msg];
        last = std::max(last, m.sent);
        // noc_update_all_tracked_messages(m.sent): retire what has been received by now
        if (!in_flight.empty() && m.sent >= in_flight.front().received)
        {
            due.clear();
            while (!in_flight.empty() && m.sent >= in_flight.front().received)
            {
                std::pop_heap(in_flight.begin(), in_flight.end(), later_received);
                due.push_back(in_flight.back());
                in_flight.pop_back();
            }
            if (due.size() > 1)
                std::sort(due.begin(), due.end(), [](const InFlight &a, const InFlight &b) {
                    return a.dest_core != b.dest_core ? a.dest_core < b.dest_core : a.seq < b.seq;
                });
            for (const InFlight &x : due)
            {
                msgs[x.msg].in_noc = false;
                noc.track(msgs[x.msg], false);
            }
        }
        if (!m.placeholder)
        {
            // schedule_handle_message  src/schedule.cpp:306-358
            const double along_route = noc.congestion(m);
            const double capacity = static_cast<double>((m.hops + 1UL) * t.noc_buffer_size);
            double blocking = 0.0;
            if (along_route > capacity)
            {
                blocking = (along_route - capacity) * noc.mean_rx;
                m.sent += blocking;
            }
            const double congestion_delay = along_route * noc.mean_rx / (static_cast<double>(m.hops) + 1.0);
            const double network_delay = std::max(m.min_hop_delay, congestion_delay);
            const double earliest = m.sent + network_delay;
            m.received = std::max(noc.core_free[m.dest_core], earliest);
            noc.core_free[m.dest_core] =
                    std::max(noc.core_free[m.dest_core] + m.processing_delay, earliest + m.processing_delay);
            m.processed = noc.core_free[m.dest_core];
            m.in_noc = true;
            in_flight.push_back({m.received, m.dest_core, seq++, ev.msg});
            std::push_heap(in_flight.begin(), in_flight.end(), later_received);
            noc.track(m, true);
            last = std::max(last, m.processed);
            if (trace != nullptr)
            {
                MessageRecord &r = recs[m.slot];
                r.sent = m.sent;
                r.received = m.received;
                r.processed = m.processed;
                r.network_delay = network_delay;
                r.blocking_delay = blocking;
                r.messages_along_route = along_route;
            }
        }
        const uint32_t c = m.src_core;
        if (head_[c] != core_begin_[c + 1])
        {
            // schedule_push_next_message  src/schedule.cpp:360-378
            Msg &next = msgs[head_[c]++];
            next.sent = m.sent + next.generation_delay;
            if (trace != nullptr) recs[next.slot].sent = next.sent;
            pq.push({next.sent, static_cast<uint32_t>(&next - msgs.data())});
            last = std::max(last, next.sent);
        }
    }
    finish_trace();
    return last + t.sync_delay;
}

} // namespace sfe
