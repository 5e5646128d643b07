// oracle/shim/coin/IpIpoptApplication.hpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Stand-in for Ipopt::IpoptApplication as used at
// mpc_ros/include/cppad/ipopt/solve.hpp:446 (construction), :527-534
// (Options()->Set*Value), :558 (Initialize) and :586 (OptimizeTNLP).
// OptimizeTNLP is implemented in oracle/shim/ip_standin.cpp on top of
// oracle/ipm.c -- NOT Ipopt; every report that uses it says so.
#ifndef ORACLE_SHIM_IPIPOPTAPPLICATION_HPP
#define ORACLE_SHIM_IPIPOPTAPPLICATION_HPP
#include "IpTNLP.hpp"
#include <map>
#include <string>

namespace Ipopt {

enum ApplicationReturnStatus {
    Solve_Succeeded = 0,
    Solved_To_Acceptable_Level = 1,
    Infeasible_Problem_Detected = 2,
    Search_Direction_Becomes_Too_Small = 3,
    Diverging_Iterates = 4,
    User_Requested_Stop = 5,
    Feasible_Point_Found = 6,
    Maximum_Iterations_Exceeded = -1,
    Restoration_Failed = -2,
    Error_In_Step_Computation = -3,
    Maximum_CpuTime_Exceeded = -4,
    Not_Enough_Degrees_Of_Freedom = -10,
    Invalid_Problem_Definition = -11,
    Invalid_Option = -12,
    Invalid_Number_Detected = -13,
    Unrecoverable_Exception = -100,
    NonIpopt_Exception_Thrown = -101,
    Insufficient_Memory = -102,
    Internal_Error = -199
};

class OptionsList : public ReferencedObject {
public:
    bool SetStringValue(const std::string &tag, const std::string &value) { str_[tag] = value; return true; }
    bool SetNumericValue(const std::string &tag, Number value) { num_[tag] = value; return true; }
    bool SetIntegerValue(const std::string &tag, Index value) { int_[tag] = value; return true; }
    bool GetNumericValue(const std::string &tag, Number &value) const
    {
        std::map<std::string, Number>::const_iterator it = num_.find(tag);
        if (it == num_.end()) return false;
        value = it->second;
        return true;
    }
    bool GetIntegerValue(const std::string &tag, Index &value) const
    {
        std::map<std::string, Index>::const_iterator it = int_.find(tag);
        if (it == int_.end()) return false;
        value = it->second;
        return true;
    }
private:
    std::map<std::string, std::string> str_;
    std::map<std::string, Number> num_;
    std::map<std::string, Index> int_;
};

class IpoptApplication : public ReferencedObject {
public:
    IpoptApplication() : options_(new OptionsList()) {}
    SmartPtr<OptionsList> Options() { return options_; }
    ApplicationReturnStatus Initialize() { return Solve_Succeeded; }
    ApplicationReturnStatus OptimizeTNLP(const SmartPtr<TNLP> &tnlp);
private:
    SmartPtr<OptionsList> options_;
};

}  // namespace Ipopt
#endif
