// include/gameai/random_generator.h -- same shape as c++/common/utils/random_generator.h:4-8
#pragma once
#include <vector>
struct IRandomGenerator
{
	virtual std::vector<int> generateUniform(int lower, int upper, int number_of_samples) = 0;
	virtual void release() = 0;
protected:
	virtual ~IRandomGenerator() {}
};
